/* fkmc.h -- C ABI of the B200-native fk_mc weight-evaluation hot path (libfkmc_b200.so).
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, returns an int status
 * (FKMC_OK == 0) and never throws.  Host pointers unless the name ends in `_dev`.
 * A context owns one GPU's device buffers, its stream and its Chebyshev tables; it is not
 * re-entrant, but different contexts may be driven from different host threads.
 *
 * Each function cites the reference interface (aeantipov/fk_mc, file:line) it replaces; the
 * binding a maintainer would add on the reference side is shown in INTEGRATION.md.
 */
#ifndef FKMC_H_
#define FKMC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fkmc_ctx fkmc_ctx;

enum fkmc_status {
    FKMC_OK = 0,
    FKMC_ERR_INVALID = 1,      /* bad argument (maps to std::logic_error in the C++ wrapper) */
    FKMC_ERR_CUDA = 2,         /* CUDA runtime error; see fkmc_last_error */
    FKMC_ERR_NO_DEVICE = 3,    /* no usable sm_100 device: there is NO CPU fallback */
    FKMC_ERR_NOCONV = 4,       /* bisection / Lanczos iteration cap hit */
    FKMC_ERR_STATE = 5         /* call sequence error (e.g. run before init) */
};

/* Lattices: include/fk_mc/lattice/hypercubic.hpp:16-95, src/lattice/hypercubic.cpp:116-203.
 * Site index is row-major with the LAST coordinate fastest (hypercubic.cpp:31-51). */
enum fkmc_lattice_kind {
    FKMC_CUBIC1D = 1,             /* fill_nearest_neighbors<1> */
    FKMC_CUBIC2D = 2,             /* fill_nearest_neighbors<2> */
    FKMC_CUBIC3D = 3,             /* fill_nearest_neighbors<3> */
    FKMC_TRIANGULAR = 4,          /* fill_triangular(t, tp) */
    FKMC_HONEYCOMB = 5,           /* intended brick wall (symmetric) */
    FKMC_HONEYCOMB_REF_LOWER = 7  /* literal fill_honeycomb as Eigen's lower-triangle solver sees it */
};

#define FKMC_MAX_W 8
#define FKMC_MAX_COND_W 32

enum fkmc_move_kind { FKMC_MOVE_FLIP = 0, FKMC_MOVE_ADDREMOVE = 1, FKMC_MOVE_RESHUFFLE = 2 };

/* ---- context ------------------------------------------------------------------------- */
/* Replaces hypercubic_lattice<D>(L) + fill_* (hypercubic.cpp:7-14,116-203).  max_batch = the
 * largest B any later call will pass (device workspaces are sized once). */
int fkmc_create(fkmc_ctx** ctx, int device, int lattice_kind, int L, double t, double tp, int max_batch);
int fkmc_destroy(fkmc_ctx* ctx);
const char* fkmc_last_error(const fkmc_ctx* ctx); /* ctx may be NULL: last creation error */
int fkmc_volume(const fkmc_ctx* ctx);             /* abstract_lattice::volume(), lattice.hpp:20 */
int fkmc_set_stream(fkmc_ctx* ctx, void* cuda_stream); /* run on a caller-owned cudaStream_t */
int fkmc_sync(fkmc_ctx* ctx);
/* dense hopping matrix, row-major H[i*N+j] = hopping_m(i,j) (abstract_lattice::hopping_m, lattice.hpp:22) */
int fkmc_hopping_dense(const fkmc_ctx* ctx, double* H);

/* ---- weight evaluators (the seam: configuration_t::calc_ed / calc_chebyshev) ------------ */
/* configuration_t::calc_hamiltonian + calc_ed(false), src/configuration.cpp:79-91,208-246.
 * f: [B][V] int32 occupations (configuration_t::f_config_).  evals: [B][N] ascending (cached_spectrum)
 * or NULL.  logZ: [B] (ed_cache::logZ).  cached_exp / cached_fermi: [B][N] or NULL. */
int fkmc_logz_ed_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta,
                         double* evals, double* logZ, double* cached_exp, double* cached_fermi);

/* calc_ed(true), src/configuration.cpp:213,216-219: evecs [B][N][N] column-major, column k <-> evals[k]. */
int fkmc_eigh_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta,
                      double* evals, double* evecs, double* logZ);

/* measure_ipr::accumulate, include/fk_mc/measures/ipr.hpp:39-56: calc_ed(true) + ipr_k = ||psi_k||_4 / ||psi_k||_2^2 for
 * every eigenstate; ipr: [B][N] (ipr_k <-> evals[k]).  The eigenvectors stay on the device. */
int fkmc_ipr_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta, double* evals, double* ipr);

/* measure_stiffness::accumulate, include/fk_mc/measures/stiffness.hpp:129-187 (cubic2d / cubic3d, current along the first
 * coordinate): calc_ed(true), mJ = V^T Jm V as a DMMA GEMM, Kubo sums on the device.  stiffness: [B]; cond: [B][n_w] optical
 * conductivity on wgrid[n_w] with Lorentzian broadening `offset` (cond_offset, fk_mc.hxx:197); n_w may be 0. */
int fkmc_stiffness_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta, double offset, int n_w,
                           const double* wgrid, double* stiffness, double* cond);

/* chebyshev_eval(max_moment, grid) + calc_chebyshev, include/fk_mc/chebyshev.hpp:21-54,
 * src/configuration.cpp:94-205.  moments: [B][M] (chebyshev_cache::moments), ab: [B][4] =
 * {e_min, e_max, a, b}, logZ: [B].  M must be even. */
int fkmc_logz_kpm_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta, int M, int G,
                          double* moments, double* ab, double* logZ);

/* The same evaluation for a configuration f that differs from a reference configuration f_ref in one or two sites (what a move
 * proposes: src/moves_chebyshev.cpp:8-24, 41-58), given the record state_ref [B][FKMC_KPM_STATE_DOUBLES] an earlier call returned for
 * f_ref: only the lattice columns within M/2 - 1 hops of the changed sites are re-evaluated (both configurations, under the scaling of
 * the record) and the moments are re-expanded for f's own e_min / e_max; same results to rounding.  state_out (may be NULL) receives
 * the record of f -- keep it with the configuration it belongs to (configuration_t::cheb_data_) and pass it when that configuration is
 * the reference.  f_ref = state_ref = NULL, more than two changed sites, a record marked invalid or a scaling that moved by more than
 * 2 % make the call a full evaluation.  Lattices outside the two-kernel 2-D path return records marked invalid.
 * Host buffers: any memory works; page-locked buffers (cudaHostAlloc / cudaHostRegister) that are reused from call to call make the copies
 * asynchronous at full link rate -- f_ref and state_ref then travel on a second stream while the e_min / e_max kernel, which only reads f,
 * already runs (the call still returns with all results in place). */
#define FKMC_KPM_STATE_DOUBLES 64
int fkmc_logz_kpm_batched_local(fkmc_ctx* ctx, const int32_t* f, const int32_t* f_ref, const double* state_ref, int B, double U,
                                double mu_c, double beta, int M, int G, double* moments, double* ab, double* logZ, double* state_out);

/* measure_energy::accumulate, src/measures/energy.cpp:6-26, from a spectrum: out [B][3] = {E_c, d2E, logZ};
 * E = E_c - mu_f*N_f + E_ff is finished by the caller. */
int fkmc_energy_from_spectrum(fkmc_ctx* ctx, const double* evals, int B, double beta, double* out3);

/* ---- stage-level entry points (used by the parity tests) --------------------------------- */
/* Householder tridiagonalisation of B dense symmetric matrices (lower triangle read),
 * A: [B][N][N] column-major.  d: [B][N], e: [B][N-1].  Eigen's tridiagonalization_inplace stage of
 * SelfAdjointEigenSolver (call site src/configuration.cpp:213). */
int fkmc_sytrd_batched(fkmc_ctx* ctx, const double* A, int N, int B, double* d, double* e);
/* Two-stage variant of the same stage (default for calc_ed): dense -> band (half-bandwidth 8, all O(N^3)
 * work on FP64 tensor cores) -> tridiagonal (bulge chasing in shared memory).
 * AB: [B][9][N] band storage, AB[d][c] = A(c+d, c). */
int fkmc_sy2sb_batched(fkmc_ctx* ctx, const double* A, int N, int B, double* AB);
int fkmc_sb2st_batched(fkmc_ctx* ctx, const double* AB, int N, int B, double* d, double* e);
/* options: "tridiag" = 1 (one-stage blocked sytrd) | 2 (two-stage, default);
 *          "kpm_generic" = 1 forces the full-lattice-vector KPM kernel (default 0: local-patch kernel when it applies)
 *          "kpm_v1" = 1 forces the single-kernel KPM path (csrc/kpm.cu) where the two-kernel 2-D path (csrc/kpm2d.cu:
 *          strip Lanczos + ring-ordered patch recursion) would apply; for cross-checks
 *          "kpm_local" = 0: Chebyshev moves of the chain engine evaluate the full trace for every proposal (default 1: only the
 *          columns near the changed sites, csrc/kpm2d.cu; takes effect at the next fkmc_chain_init); "kpm_rebase_sweeps" = n: sweeps
 *          between recomputations of the per-chain trace sums from scratch (default 16, 0 = never)
 *          "band_path" = 0 sends eigenvalue-only solves through the dense N^3 reduction (default 1: lattices whose matrix has
 *          half-bandwidth <= 64 after folding the slow coordinate start from the band, csrc/sb2sb.cu); "band_min" = n: smallest N
 *          served by the band path (default 256)
 *          "sy2sb_tiled_min" = n: smallest N served by the tiled dense->band kernel (per context, default 256; below it the
 *          column-major small-matrix kernel runs); for cross-checks
 *          "kpm_generic_schedule" = 1 keeps the moments kernel of csrc/kpm2d.cu on its run-time slot schedule (default 0: the
 *          compile-time schedule where one is known, i.e. cubic2d with a single hopping constant); for cross-checks
 *          "cuda_graph" = 0 launches the kernels of every Metropolis step one by one (default 1: fkmc_chain_run_sweeps captures a
 *          step once and replays the graph; event profiling and trace recording also run without the graph)
 *          "eigvec_v1" = 1 back-transforms the eigenvectors reflector by reflector (default 0: compact-WY groups of 32 on DMMA)
 *          "lanczos_max_steps" = n > 0 lowers the Lanczos step cap of the KPM kernels (default 0: 384); the tests use it to force
 *          FKMC_ERR_NOCONV */
int fkmc_set_option(fkmc_ctx* ctx, const char* name, int value);
/* eigenvalues (ascending) of B symmetric tridiagonals by Sturm bisection */
int fkmc_tridiag_eigvals_batched(fkmc_ctx* ctx, const double* d, const double* e, int N, int B, double* evals);
/* eigenvalues (ascending) of diag(lam) + rho z z^T for B problems: lam [B][N] ascending, z [B][N], rho [B] -> lam_new [B][N].
 * The rank-one secular solver behind chain parameter fast_update (csrc/secular.cu; the step benchmark/fast_update.cpp times). */
int fkmc_secular_update_batched(fkmc_ctx* ctx, const double* lam, const double* z, const double* rho, int N, int B, double* lam_new);
/* device std::mt19937 + libstdc++ distributions: mode 0 raw words, 1 uniform_int(0,V-1), 2 uniform_real(0,1) */
int fkmc_rng_stream(fkmc_ctx* ctx, int64_t seed, int mode, int V, int count, double* out);

/* ---- device-resident Markov chains (mc_metropolis + moves + measures) -------------------- */
/* Mirrors fk_mc<L>::define_parameters (include/fk_mc/fk_mc.hxx:177-207) and
 * mc_metropolis::define_parameters (src/mc_metropolis.cpp:11-19). */
typedef struct fkmc_chain_params {
    double beta, U, mu_c, mu_f;
    double mc_flip, mc_add_remove, mc_reshuffle; /* move weights; a move is registered iff weight > eps */
    int32_t cheb_moves;                          /* 0: exact moves (src/moves.cpp), 1: src/moves_chebyshev.cpp */
    double cheb_prefactor;                       /* M = even(int(ln N * prefactor)), G = max(2M,10) */
    int64_t seed;                                /* chain c uses std::mt19937(seed + chain0 + c), mc_metropolis.cpp:25 */
    int32_t chain0;                              /* global id of this context's first chain (rank offset) */
    int32_t nf_start;                            /* randomize_f(rng, nf_start); the exec passes V/2 */
    int32_t sweep_len;                           /* proposals per sweep */
    int32_t ntherm_sweeps;                       /* sweeps before measuring starts */
    int32_t measure_energy;                      /* energy/spectrum measures (exact calc_ed per measured sweep) */
    int32_t record_trace;                        /* keep per-step (site, weight, u, accepted) for parity tests */
    int32_t max_sweeps;                          /* capacity of the series / trace / history buffers */
    int32_t measure_history;                     /* fk_mc.hxx:185: spectrum_history (when an exact spectrum is measured) and focc_history */
    int32_t measure_ipr;                         /* fk_mc.hxx:195: measure_ipr -> ipr_history (calc_ed(true) per measured sweep) */
    int32_t n_W;                                 /* 1-D lattices: f-f interaction W[0..n_W) (config_params::W, configuration.hpp:15-19); */
    double W[FKMC_MAX_W];                        /* ignored for D >= 2 where calc_ff_energy() == 0 (configuration.cpp:62) */
    int32_t measure_eigenfunctions;              /* fk_mc.hxx:196: eigenfunctions_history, N x N eigenvector matrix per chain and measured sweep
                                                    (max_sweeps * n_chains * N^2 doubles of device memory) */
    int32_t measure_stiffness;                   /* fk_mc.hxx:198 (registration commented out at HEAD, fk_mc.hxx:101-105): measure_stiffness per measured
                                                    sweep -> stiffness series and cond_history on cond_wgrid; cubic2d / cubic3d only */
    int32_t n_cond_w;                            /* number of conductivity frequencies (<= FKMC_MAX_COND_W; the exec's wgrid_conductivity) */
    double cond_offset;                          /* Lorentzian broadening (fk_mc.hxx:197, default 0.05) */
    double cond_wgrid[FKMC_MAX_COND_W];
    int32_t fast_update;                         /* exact moves only: 1 = re-weight add_remove / flip proposals through rank-one secular
                                                    updates of the tracked eigen-decomposition (O(N^2) per proposal, an N^3 eigenvector
                                                    update only on accept; benchmark/fast_update.cpp) instead of a fresh
                                                    eigensolve per proposal.  Same spectra (1e-13) and accept sequence.  Needs N <= 1024,
                                                    mc_reshuffle == 0 and 3 N^2 doubles of device memory per chain. */
    int32_t fu_refresh_sweeps;                   /* fast_update: full re-diagonalisation + consistency check every this many sweeps
                                                    (0: default 64) */
} fkmc_chain_params;

int fkmc_chain_init(fkmc_ctx* ctx, int n_chains, const fkmc_chain_params* p);
/* mc_metropolis::update + measure, n_sweeps times (src/mc_metropolis.cpp:34-61).  Synchronises once at the end of the call
 * and returns FKMC_ERR_NOCONV when any evaluation of these sweeps hit the Lanczos / bisection iteration cap. */
int fkmc_chain_run_sweeps(fkmc_ctx* ctx, int n_sweeps);
/* series: [n_measured][n_chains] each (observables_t::energies, d2energies, c_energies, fk_mc.hpp:13-34); any may be NULL */
int fkmc_chain_get_series(fkmc_ctx* ctx, int* n_measured, double* energies, double* d2energies, double* c_energies,
                          int32_t* nf);
/* f-sector series of the measured sweeps, [n_measured][n_chains] each (measure_nf0pi, include/fk_mc/measures/fsusc0pi.hpp:36-46:
 * observables_t::nf0 = sum_i f_i and nfpi = |sum_i (-1)^(x+y+..) f_i|); either may be NULL */
int fkmc_chain_get_fsector(fkmc_ctx* ctx, int* n_measured, int32_t* nf0, int32_t* nfpi);
/* state: f [n_chains][V], logZ [n_chains], naccept [n_chains] */
int fkmc_chain_get_state(fkmc_ctx* ctx, int32_t* f, double* logZ, int64_t* naccept, double* spectrum);
/* trace: [n_steps][n_chains] each; n_steps = sweeps run * sweep_len */
int fkmc_chain_get_trace(fkmc_ctx* ctx, int* n_steps, int32_t* move, int32_t* site_a, int32_t* site_b, int32_t* accepted,
                         double* weight, double* u, double* logz_new);
/* per-sweep histories of the measured sweeps (any pointer may be NULL; FKMC_ERR_STATE when that measure is off):
 *   spectrum_mean    [n_chains][N]              measure_spectrum, src/measures/spectrum.cpp:13-21 (running mean per chain)
 *   spectrum_history [n_measured][n_chains][N]  measure_spectrum_history, src/measures/spectrum_history.cpp:13-19
 *   focc_history     [n_measured][n_chains][V]  measure_focc, src/measures/focc_history.cpp:7-12
 *   ipr_history      [n_measured][n_chains][N]  measure_ipr, include/fk_mc/measures/ipr.hpp:39-56 */
int fkmc_chain_get_history(fkmc_ctx* ctx, int* n_measured, double* spectrum_mean, double* spectrum_history, int32_t* focc_history,
                           double* ipr_history);
/* measure_eigenfunctions (src/measures/eigenfunctions.cpp:12-18): evecs [n_measured][n_chains][N][N], each matrix column-major like
 * ed_cache::cached_evecs (column k <-> eigenvalue k of the spectrum history); needs chain parameter measure_eigenfunctions */
int fkmc_chain_get_eigenfunctions(fkmc_ctx* ctx, int* n_measured, double* evecs);
/* measure_stiffness series (include/fk_mc/measures/stiffness.hpp:129-187): stiffness [n_measured][n_chains] (observables_t::stiffness) and
 * cond [n_measured][n_chains][n_cond_w] (observables_t::cond_history, frequency-major there); needs chain parameter measure_stiffness */
int fkmc_chain_get_stiffness(fkmc_ctx* ctx, int* n_measured, double* stiffness, double* cond);
/* measure_ipr on the chains' current configurations: evals [n_chains][N] (or NULL), ipr [n_chains][N] */
int fkmc_chain_ipr(fkmc_ctx* ctx, double* evals, double* ipr);
/* device pointers to the series (for the end-of-run NCCL gather): energies, d2energies, c_energies as
 * [max_sweeps][n_chains] doubles */
int fkmc_chain_series_dev(fkmc_ctx* ctx, void** energies, void** d2energies, void** c_energies, int* ld);

/* ---- end-of-run collective over NCCL (NVLink / NVSwitch), one context per rank ------------------ */
/* Replaces the root-0 reduce/gather of measure_energy::collect_results (src/measures/energy.cpp:32-47).  libnccl is loaded at run
 * time (dlopen "libnccl.so.2"; FKMC_NCCL_LIB overrides), so single-GPU use needs no NCCL.
 * fkmc_nccl_unique_id: rank 0 fills a 128-byte ncclUniqueId which the caller distributes to the other ranks by any host channel.
 * fkmc_comm_init: every rank joins the communicator (collective call).
 * fkmc_gather_series: collective.  All ranks run the same number of chains and measured sweeps.  The three energy series are
 *   all-gathered from the chain engine's device buffers and reordered on the device to [n_measured][nranks * n_chains] in global
 *   chain order (chain id = the reference's MPI rank).  Host pointers (any may be NULL) receive copies; *out_dev (may be NULL) is
 *   set to the device buffer holding energies | d2energies | c_energies back to back.  Without a communicator: the local series. */
int fkmc_nccl_unique_id(void* id128);
int fkmc_comm_init(fkmc_ctx* ctx, const void* id128, int nranks, int rank);
int fkmc_comm_destroy(fkmc_ctx* ctx);
int fkmc_gather_series(fkmc_ctx* ctx, int* n_measured, int* total_chains, double* energies, double* d2energies, double* c_energies,
                       void** out_dev);

/* ---- instrumentation -------------------------------------------------------------------- */
/* Lanczos steps each of the last B KPM evaluations needed for e_min / e_max (the reference's ARPACK iteration count analogue) */
int fkmc_kpm_last_steps(fkmc_ctx* ctx, int B, int32_t* steps);
/* number of kernels this context has launched since creation */
int64_t fkmc_launch_count(const fkmc_ctx* ctx);
/* CUDA-event timing on the context's stream: begin/end a region, read milliseconds */
int fkmc_timer_begin(fkmc_ctx* ctx);
int fkmc_timer_end(fkmc_ctx* ctx, float* ms);
/* accumulated device time of one kernel family since the last reset (events around each launch when enabled) */
int fkmc_profile_enable(fkmc_ctx* ctx, int on);
int fkmc_profile_get(fkmc_ctx* ctx, const char* family, double* total_ms, int64_t* launches);
int fkmc_profile_reset(fkmc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* FKMC_H_ */
