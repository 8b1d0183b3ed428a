// fk_mc CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement of the reference's Metropolis weight-evaluation path
// (aeantipov/fk_mc).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product (fk_mc_b200/) never does.
//
// The reference itself cannot be compiled in this image (Eigen, Boost.MPI, ALPSCore, ARPACK,
// FFTW, HDF5 are absent), so there is no oracle/_ref binary.  What pins this restatement:
//   * the reference tests' golden values (Bessel moments test/fast_update_test.cpp:62-65,
//     logZ identity :78, stiffness goldens test/stiffness_test.cpp:60-63, binning / jackknife
//     goldens test/binning_test.cpp:60-63, test/jackknife_test.cpp:56-134, E_ff known answer
//     test/config_test.cpp:46-72, hopping sum test/honeycomb_test.cpp:26),
//   * analytic spectra (free lattice, checkerboard) and LAPACK (scipy) on seeded configurations,
//     committed as fixtures under tests/golden/ together with the generating script.
// Eigenvalue parity itself is "unpinned" by the reference (no reference test asserts an
// eigenvalue; Eigen/ARPACK versions are not pinned) -- see DESIGN.md.
//
// Third-party arithmetic restated here (absent from /root/reference): Eigen
// SelfAdjointEigenSolver (call site src/configuration.cpp:213), ARPACK dsaupd nev=1 SA/LA (call
// sites src/configuration.cpp:99-100).  libstdc++ <random> is used directly (same library the
// reference links: include/fk_mc/common.hpp:18).
#pragma once
#include <cstddef>
#include <cstdint>
#include <random>
#include <utility>
#include <vector>

namespace orc {

typedef std::mt19937 random_generator;  // include/fk_mc/common.hpp:18

enum lattice_kind {
    CUBIC1D = 1,
    CUBIC2D = 2,
    CUBIC3D = 3,
    TRIANGULAR = 4,
    HONEYCOMB = 5,           // intended brick-wall: A <=> (x+y) even hops up, B hops down (symmetric)
    HONEYCOMB_REF = 6,       // literal fill_honeycomb (non-symmetric)
    HONEYCOMB_REF_LOWER = 7  // literal matrix, lower triangle mirrored (what Eigen's dense solver sees)
};

// Hopping matrix in "insertion" form: rows[i] = list of (j, value) with hopping_m(i,j) = value.
struct lattice {
    int kind = 0, ndim = 0, L = 0, N = 0;
    std::vector<std::vector<std::pair<int, double>>> rows;
    double hop(int i, int j) const;  // 0 when absent
    std::vector<int> index_to_pos(int index) const;
    int pos_to_index(const std::vector<int>& pos) const;
};
lattice make_lattice(int kind, int L, double t, double tp);

// ---- configuration (src/configuration.cpp) ----
void randomize_f(random_generator& rnd, int V, size_t nf, std::vector<int>& f);
double calc_ff_energy(int ndim, const std::vector<int>& f, const std::vector<double>& W);
// dense H, column-major n*n, lower triangle meaningful (H(r,c) r>=c = hopping(r,c) + diag)
void dense_hamiltonian(const lattice& lat, const std::vector<int>& f, double U, double mu_c, std::vector<double>& H);

// ---- Eigen::SelfAdjointEigenSolver restatement ----
// A: column-major n*n, only the lower triangle is read.  evals ascending.  evecs (optional,
// column-major, column k <-> evals[k]).  Returns 0, or 1 when the QR iteration hit 30*n sweeps.
int eigh_lower(int n, const double* A, double* evals, double* evecs);
// tridiagonal produced by the Householder stage (for diagnostics)
void tridiagonalize_lower(int n, std::vector<double>& A, std::vector<double>& diag, std::vector<double>& sub,
                          std::vector<double>* hcoeffs);
int tridiag_ql_implicit(int n, double* diag, double* sub, double* Q /*nullable, col-major n*n*/);

struct ed_result {
    std::vector<double> spectrum, cached_exp, cached_fermi, evecs;
    double logZ = 0;
};
void calc_ed(const lattice& lat, const std::vector<int>& f, double U, double mu_c, double beta, bool evecs, ed_result& out);
double logz_from_spectrum(const std::vector<double>& spectrum, double beta, std::vector<double>* cexp, std::vector<double>* cfermi);

// ---- chebyshev (include/fk_mc/chebyshev.hpp) ----
struct chebyshev_eval {
    int M, G;
    std::vector<double> angle_grid, lobatto_grid, chebt;  // chebt[k*G + i]
    chebyshev_eval(int max_moment, int grid_size);
    double moment(const std::vector<double>& vals, int order) const;
    template <class F> double moment_f(F&& op, int order) const {
        std::vector<double> vals(G);
        for (int i = 0; i < G; ++i) vals[i] = op(lobatto_grid[i]);
        return moment(vals, order);
    }
};
void cheb_sizes(int msize, double prefactor, int& M, int& G);  // include/fk_mc/fk_mc.hxx:60-63

struct cheb_result {
    double e_min = 0, e_max = 0, a = 0, b = 0, logZ = 0;
    std::vector<double> moments;
    int lanczos_steps = 0;
};
// emode: 0 = extremal eigenvalues from the dense spectrum, 1 = Lanczos (ARPACK stand-in)
void calc_chebyshev(const lattice& lat, const std::vector<int>& f, double U, double mu_c, double beta,
                    const chebyshev_eval& cheb, int emode, bool prune, cheb_result& out);
void lanczos_extremal(const lattice& lat, const std::vector<double>& diag, double& e_min, double& e_max, int& steps);

// ---- measures ----
void measure_energy(const ed_result& ed, double beta, double mu_f, int nf, double eff, double& e, double& d2e, double& ec);
void measure_ipr(int n, const std::vector<double>& evecs, std::vector<double>& ipr);
double measure_stiffness(const lattice& lat, const ed_result& ed, double beta, double offset,
                         const std::vector<double>* wgrid, std::vector<double>* cond);

// ---- statistics (include/fk_mc/binning.hpp, jackknife.hpp) ----
struct bin_stats { double n, mean, var, err; };
bin_stats calc_stats(const std::vector<double>& x);
std::vector<double> bin_once(const std::vector<double>& x);                       // pairwise averaging, depth 1
std::vector<bin_stats> accumulate_binning(const std::vector<double>& x, int max_depth);
double calc_cor_length(const std::vector<bin_stats>& b, int level);

// ---- Monte Carlo driver (src/mc_metropolis.cpp + src/moves*.cpp + fk_mc.hxx) ----
struct mc_params {
    int kind, L;
    double t, tp, beta, U, mu_c, mu_f;
    double mc_flip, mc_add_remove, mc_reshuffle;
    int cheb_moves;
    double cheb_prefactor;
    int emode;  // extremal-eigenvalue mode for the Chebyshev path
    long seed;
    int nf_start;
    int nsweeps, sweep_len, ntherm_sweeps;
    int measure_energy;  // register energy/spectrum measures (exact calc_ed per sweep)
    int measure_ipr;
    int n_W;        // 1-D lattices: f-f interaction W[0..n_W) (config_params::W); calc_ff_energy() is 0 for D >= 2
    double W[8];
};
struct mc_trace {
    std::vector<int> move, site_a, site_b, accepted;
    std::vector<double> weight, u, logz_new;
};
struct mc_result {
    std::vector<double> energies, d2energies, c_energies, spectrum_avg;
    std::vector<std::vector<double>> ipr_history;  // [measurement][state]
    std::vector<std::vector<double>> spectrum_history;  // [measurement][index]   (src/measures/spectrum_history.cpp:13-19)
    std::vector<std::vector<int>> focc_history;         // [measurement][site]    (src/measures/focc_history.cpp:7-12)
    std::vector<int> f_final, nf_series;
    long naccept = 0;
    double logz_final = 0;
};
void mc_run(const mc_params& p, int rank, mc_result& res, mc_trace* trace);

}  // namespace orc
