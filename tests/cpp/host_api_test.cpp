// Drives the C++ host mirror (include/fk_mc_b200/fk_mc.hpp) the way prog/fk_mc_exec.cpp + fk_mc.hxx:35-125 drive the
// reference: lattice -> configuration_t -> randomize_f -> register moves / measures -> run.  Prints the observables so that
// tests/test_gpu_parity.py can compare them with the CPU oracle.  usage: host_api_test L beta U cheb(0|1) mc_flip nsweeps seed rank
#include <cstdio>
#include <cstdlib>

#include "fk_mc_b200/fk_mc.hpp"

using namespace fk;

int main(int argc, char** argv) {
    if (argc < 9) return 2;
    const int L = atoi(argv[1]);
    const double beta = atof(argv[2]), U = atof(argv[3]);
    const bool cheb_move = atoi(argv[4]) != 0;
    const double mc_flip = atof(argv[5]);
    const int nsweeps = atoi(argv[6]);
    const long seed = atol(argv[7]);
    const int rank = atoi(argv[8]);
    try {
        typedef hypercubic_lattice<2> lattice_t;
        lattice_t lattice(L);
        fill_nearest_neighbors(lattice, 1.0);
        alps::mc_metropolis mc(seed, rank, nsweeps, /*sweep_len=*/16, /*ntherm_sweeps=*/1);
        configuration_t config(lattice, beta, U, U / 2, U / 2);
        config.randomize_f(mc.rng(), lattice.volume() / 2);  // fk_mc.hxx:46, fk_mc_exec.cpp:124
        config.calc_hamiltonian();
        int cheb_size = int(std::log(lattice.msize()) * 2.2);  // fk_mc.hxx:60-63
        cheb_size += cheb_size % 2;
        chebyshev::chebyshev_eval cheb(cheb_size, std::max(cheb_size * 2, 10));
        if (mc_flip > std::numeric_limits<double>::epsilon()) {
            if (!cheb_move) mc.add_move(move_flip(beta, config, mc.rng()), "flip", mc_flip);
            else mc.add_move(chebyshev::move_flip(beta, config, cheb, mc.rng()), "flip", mc_flip);
        }
        if (!cheb_move) mc.add_move(move_addremove(beta, config, mc.rng()), "add_remove", 1.0);
        else mc.add_move(chebyshev::move_addremove(beta, config, cheb, mc.rng()), "add_remove", 1.0);
        std::vector<double> energies, d2energies, c_energies;
        mc.add_measure(measure_energy(beta, config, energies, d2energies, c_energies), "energy");
        mc.run();
        printf("naccept %ld\n", mc.naccept());
        printf("energies");
        for (double e : energies) printf(" %.17g", e);
        printf("\nd2energies");
        for (double e : d2energies) printf(" %.17g", e);
        printf("\nf");
        for (int f : config.f_config_) printf(" %d", f);
        printf("\n");
        // error behaviour: mismatched parameters throw std::logic_error like configuration.cpp:38
        configuration_t other(lattice, beta + 1, U, U / 2, U / 2);
        bool threw = false;
        try { other = config; } catch (std::logic_error&) { threw = true; }
        printf("mismatch_throws %d\n", int(threw));
        bool threw2 = false;
        try { hypercubic_lattice<2> odd(7); fill_honeycomb(odd, 1.0); } catch (std::logic_error&) { threw2 = true; }
        printf("honeycomb_odd_throws %d\n", int(threw2));
    } catch (std::exception& e) {
        printf("exception %s\n", e.what());
        return 1;
    }
    return 0;
}
