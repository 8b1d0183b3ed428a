// Drives the C++ host mirror (include/fk_mc_b200/fk_mc.hpp) the way prog/fk_mc_exec.cpp drives the reference:
//   parameters_t p; fk_mc<lattice_t>::define_parameters(p); ...; fk_mc<lattice_t> mc(p, rank); mc.initialize(lattice); mc.run();
// (prog/fk_mc_exec.cpp:43,121-151, include/fk_mc/fk_mc.hxx:35-125,177-207), then the same run as a batch of chains resident on the
// GPU (fk_mc::run_batched -> fkmc_chain_*).  Prints the observables so that tests/test_gpu_parity.py can compare them with the
// CPU oracle.  usage: host_api_test L beta U cheb(0|1) mc_flip nsweeps seed rank
#include <cstdio>
#include <cstdlib>

#include "fk_mc_b200/fk_mc.hpp"

using namespace fk;

static void print_series(const char* name, const std::vector<double>& v) {
    printf("%s", name);
    for (double e : v) printf(" %.17g", e);
    printf("\n");
}

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
        typedef fk_mc<lattice_t> qmc_t;
        lattice_t lattice(L);
        fill_nearest_neighbors(lattice, 1.0, /*device=*/0, /*max_batch=*/2);
        parameters_t p;
        p["seed"] = seed;                       // command-line values are set before define_parameters, as alps::params does
        qmc_t::define_parameters(p);
        if (int(p["sweep_len"]) != 16 || int(p["ntherm_sweeps"]) != 1 || double(p["cheb_prefactor"]) != 2.2 || double(p["mc_add_remove"]) != 1.0) {
            printf("exception defaults differ from fk_mc.hxx:177-207 / mc_metropolis.cpp:11-19\n");
            return 1;
        }
        p["beta"] = beta; p["U"] = U; p["mu_c"] = U / 2; p["mu_f"] = U / 2;
        p["mc_flip"] = mc_flip; p["cheb_moves"] = cheb_move; p["nsweeps"] = nsweeps;
        p["Nf_start"] = int(lattice.volume() / 2);   // fk_mc_exec.cpp:124
        p["measure_ipr"] = cheb_move;                // Chebyshev moves only measure the energy when an exact spectrum is asked for (fk_mc.hxx:110-118)
        qmc_t mc(p, rank);
        mc.initialize(lattice, true);
        mc.run();
        printf("naccept %ld\n", mc.naccept());
        print_series("energies", mc.observables.energies);
        print_series("d2energies", mc.observables.d2energies);
        print_series("spectrum_mean", mc.observables.spectrum);
        printf("history_shape %zu %zu %zu\n", mc.observables.spectrum_history.size(), mc.observables.spectrum_history.empty() ? 0 : mc.observables.spectrum_history[0].size(),
               mc.observables.focc_history.size());
        printf("f");
        for (int f : mc.config().f_config_) printf(" %d", f);
        printf("\n");
        // the batched product path: ranks rank, rank + 1 as two chains resident on the GPU (dense moves re-weighted by rank-one secular
        // updates: same results as one eigensolve per proposal)
        p["fast_update"] = !cheb_move;
        p["measure_stiffness"] = true;
        qmc_t mcb(p, rank);
        mcb.initialize(lattice, true, {0.0, 0.5});
        auto obs = mcb.run_batched(lattice, 2);
        printf("b_stiffness_shape %zu %zu %zu\n", obs[0].stiffness.size(), obs[0].cond_history.size(), obs[0].cond_history.empty() ? 0 : obs[0].cond_history[0].size());
        print_series("b0_stiffness", obs[0].stiffness);
        for (int c = 0; c < 2; ++c) {
            printf("b%d_naccept %ld\n", c, long(mcb.batched_naccept()[c]));
            print_series(c ? "b1_energies" : "b0_energies", obs[c].energies);
            print_series(c ? "b1_spectrum_mean" : "b0_spectrum_mean", obs[c].spectrum);
            printf("b%d_f", c);
            for (size_t i = 0; i < lattice.volume(); ++i) printf(" %d", mcb.batched_f_config()[c * lattice.volume() + i]);
            printf("\n");
        }
        printf("b_history_shape %zu %zu %zu\n", obs[0].spectrum_history.size(), obs[0].spectrum_history.empty() ? 0 : obs[0].spectrum_history[0].size(),
               obs[0].focc_history.size());
        // error behaviour: mismatched parameters throw std::logic_error like configuration.cpp:38
        configuration_t other(lattice, beta + 1, U, U / 2, U / 2);
        bool threw = false;
        try { other = mc.config(); } catch (std::logic_error&) { threw = true; }
        printf("mismatch_throws %d\n", int(threw));
        bool threw2 = false;
        try { hypercubic_lattice<2> odd(7); fill_honeycomb(odd, 1.0); } catch (std::logic_error&) { threw2 = true; }
        printf("honeycomb_odd_throws %d\n", int(threw2));
        bool threw3 = false;   // no registered moves: mc_metropolis.cpp:35-38
        try {
            parameters_t q = p;
            q["mc_flip"] = 0.0; q["mc_add_remove"] = 0.0;
            qmc_t none(q, 0);
            none.initialize(lattice, true);
            none.update();
        } catch (std::logic_error&) { threw3 = true; }
        printf("no_moves_throws %d\n", int(threw3));
    } catch (std::exception& e) {
        printf("exception %s\n", e.what());
        return 1;
    }
    return 0;
}
