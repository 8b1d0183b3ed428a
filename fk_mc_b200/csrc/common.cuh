// Internal declarations shared by the sm_100a kernels of libfkmc_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/fkmc.h"

#define FKMC_MAX_Z 8        // max neighbours per site (triangular: 6)
#define FKMC_SYTRD_NB 32    // panel width of the blocked tridiagonalisation
#define FKMC_KPM_STATE 64   // doubles per chain of the local KPM scheme (kpm2d.cu)
#define FKMC_MAX_HALF 16    // KPM: M/2 <= 16

struct fkmc_profile_entry {
    double total_ms = 0;
    int64_t launches = 0;
};

struct fkmc_pending_event {
    const char* name;
    cudaEvent_t a, b;
};

struct fkmc_chain_state {
    bool active = false;
    fkmc_chain_params p{};
    int n_chains = 0, M = 0, G = 0;
    int n_moves = 0;
    int move_kind[3] = {0, 0, 0};
    double move_cp[3] = {0, 0, 0};  // cumulative probabilities (std::discrete_distribution)
    long sweeps_done = 0, measured = 0;
    // one Metropolis step (propose -> evaluate -> accept [-> eigenvector update]) captured as a CUDA graph and replayed sweep_len times per sweep
    cudaGraphExec_t step_graph = nullptr;
    int step_graph_nodes = 0;
    bool step_graph_failed = false;
    // device buffers
    uint32_t* mt = nullptr;     // [n_chains][625] state + index
    int32_t* f_cur = nullptr;   // [n_chains][V]
    int32_t* f_prop = nullptr;  // [n_chains][V]
    double* logz_cur = nullptr;
    double* logz_prop = nullptr;
    double *ks_cur = nullptr, *ks_prop = nullptr;   // [n_chains][FKMC_KPM_STATE] local KPM scheme (kpm2d.cu)
    double* spec[2] = {nullptr, nullptr};  // [n_chains][N] double-buffered spectrum (exact moves)
    int32_t* cur_slot = nullptr;           // [n_chains] which of spec[] is current
    int32_t* prop_move = nullptr;          // [n_chains] move kind of the pending proposal (-1: early-out, weight 0)
    int32_t* prop_a = nullptr;
    int32_t* prop_b = nullptr;
    int64_t* naccept = nullptr;
    int32_t *nf_cur = nullptr, *nf_prop = nullptr;  // [n_chains] occupied sites of the current / proposed configuration
    int32_t* prop_slot = nullptr;                   // [n_chains] spectrum slot the pending proposal is written to
    double *ec_cur = nullptr, *d2_cur = nullptr;    // [n_chains] E_c, d2E of the current configuration (exact moves)
    double *eff_cur = nullptr, *eff_prop = nullptr; // [n_chains] calc_ff_energy() of the current / proposed configuration (1-D, W != {})
    double* d_W = nullptr;                          // [n_W] f-f interaction
    // per-sweep histories (measure_spectrum, measure_spectrum_history, measure_focc, measure_ipr)
    double* spec_mean = nullptr;     // [n_chains][N] running mean of the sorted spectrum
    double* spec_hist = nullptr;     // [max_sweeps][n_chains][N]
    int32_t* focc_hist = nullptr;    // [max_sweeps][n_chains][V]
    double* ipr_hist = nullptr;      // [max_sweeps][n_chains][N]
    double* s_stiff = nullptr;       // [max_sweeps][n_chains] stiffness series
    double* cond_hist = nullptr;     // [max_sweeps][n_chains][n_cond_w] optical conductivity
    double* d_cond_w = nullptr;      // [n_cond_w] frequency grid
    double* eig_hist = nullptr;      // [max_sweeps][n_chains][N][N] eigenvector-major (== column-major matrix)
    double* ipr_evals = nullptr;     // [n_chains][N] spectrum of the eigenvector solve of the last measured sweep
    long spec_count = 0;             // measurements folded into spec_mean
    // fast update of the dense moves (secular.cu): eigenvectors of the current configurations and per-stage root data
    double* fu_vt = nullptr;         // [2][n_chains][N][N] site-major eigenvectors, double-buffered
    double* fu_q = nullptr;          // [n_chains][N][N] right factor Q of the pending eigenvector update
    int32_t* fu_vslot = nullptr;     // [n_chains]
    double *fu_poles = nullptr, *fu_mu = nullptr, *fu_zhat = nullptr, *fu_inrm = nullptr;  // [2][n_chains][N]
    int32_t* fu_org = nullptr;       // [2][n_chains][N]
    int32_t* fu_nstage = nullptr;    // [n_chains]
    double* fu_rho = nullptr;        // [2][n_chains]
    int32_t* fu_acc = nullptr;       // [n_chains] accept flag of the last step
    double* fu_fresh = nullptr;      // [n_chains][N] spectrum of the last refresh
    double* fu_maxdev = nullptr;     // [n_chains] relative deviation tracked vs fresh spectrum at the last refresh
    double *s_energy = nullptr, *s_d2energy = nullptr, *s_cenergy = nullptr;  // [max_sweeps][n_chains]
    int32_t* s_nf = nullptr;
    int32_t* s_nfpi = nullptr;  // [max_sweeps][n_chains] |n_f(q = pi)| = |sum_i (-1)^(x+y+..) f_i|
    // trace [max_steps][n_chains]
    int32_t *t_move = nullptr, *t_a = nullptr, *t_b = nullptr, *t_acc = nullptr;
    double *t_w = nullptr, *t_u = nullptr, *t_lz = nullptr;
};

struct fkmc_ctx {
    int device = 0;
    int kind = 0, ndim = 0, L = 0, N = 0, Z = 0;
    double t = 1, tp = 1;
    int max_batch = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int64_t launches = 0;
    int num_sms = 0;
    size_t smem_optin = 0;

    // lattice tables (host + device).  nbr_idx[z*N + i] = z-th neighbour of site i (N = "none": the
    // zero slot), nbr_val[z*N + i] = hopping_m(i, nbr) (0 for padding).
    std::vector<int> h_nbr_idx;
    std::vector<double> h_nbr_val;
    int* d_nbr_idx = nullptr;
    double* d_nbr_val = nullptr;

    // workspaces sized for max_batch
    double* d_A = nullptr;      // [max_batch][N][N] dense Hamiltonians (column-major, lower triangle live)
    double* d_W = nullptr;      // [max_batch][N][NB] panel workspace
    // band path (sb2sb.cu): folded site ordering with half-bandwidth band_bw <= 64 (0: not available), static hopping tiles, work band
    int band_bw = 0, band_path = 1, band_min = 256;
    double* d_band0 = nullptr;
    int* d_band_perm = nullptr;
    double* d_band = nullptr;   // [max_batch][fkmc_band_stride(N)]
    double* d_AB = nullptr;     // [max_batch][9][N] band storage of the two-stage reduction
    double* d_s1_scratch = nullptr;  // sy2sb: fragment-ordered panel records, one slot per matrix
    size_t s1_scratch_cap = 0;       // doubles
    unsigned char* d_tile_mask = nullptr;  // tiled layout: 1 for the tiles of H that can hold a non-zero entry
    int nsmid = 0;                   // %nsmid of the device (scratch slots of sy2sb)
    int tridiag_mode = 2;       // 1: one-stage blocked sytrd, 2: sy2sb + sb2st
    int sb2st_warps = 0;        // 0: automatic
    int tiled_min = 256;        // smallest N served by the tiled dense->band kernel ("sy2sb_tiled_min")
    int lanczos_cap = 0;        // > 0: Lanczos step cap of the KPM kernels ("lanczos_max_steps"; tests force non-convergence with it)
    int use_graphs = 1;         // 0: launch every kernel of a Metropolis step individually ("cuda_graph" option)
    int eigvec_v1 = 0;          // 1: per-reflector back-transformation kernel instead of the blocked DMMA one (cross-checks)
    int kpm_force_generic = 0;  // 1: always use the full-lattice-vector KPM kernel (for cross-checks)
    int kpm_force_v1 = 0;       // 1: single-kernel KPM (kpm.cu) even where the two-kernel 2-D path (kpm2d.cu) applies
    int kpm2_H = 0;             // radius of the cached patch tables of kpm2d.cu
    unsigned long long kpm2_sched = 0;        // slots per step of the cached tables, four bits per step
    int kpm_no_sched = 0;                     // 1: never use the schedule-specialised moments kernel (cross-checks)
    int kpm2_S = 0, kpm2_PV = 0, kpm2_P = 0;  // slots per lane, cells per buffer, row pitch of the patch array
    int* d_kpm2_tabi = nullptr;
    double* d_kpm2_h1 = nullptr;
    int* d_kpm2_off = nullptr;
    unsigned short* d_kpm2_nb = nullptr;
    double* d_kpm2_part = nullptr;   // [max_batch][2][3][FKMC_MAX_HALF+1] per-CTA partial traces of the moments kernel
    int* d_kpm2_arrived = nullptr;   // [max_batch]
    int* d_kpm2_order = nullptr;     // [max_batch] launch order of the Lanczos CTAs (longest expected run first)
    double* d_d = nullptr;      // [max_batch][N]
    double* d_e = nullptr;      // [max_batch][N]
    double* d_tau = nullptr;    // [max_batch][N]
    double* d_evals = nullptr;  // [max_batch][N]
    double* d_out = nullptr;    // [max_batch][8] per-matrix scalars (logZ, E_c, d2E, ...)
    int32_t* d_f = nullptr;     // [max_batch][N] staging for host f
    int* d_flag = nullptr;      // non-convergence flag
    double* d_moments = nullptr;  // [max_batch][2*FKMC_MAX_HALF]
    double* d_ab = nullptr;       // [max_batch][4]
    // local KPM re-evaluation (kpm2d.cu): set by the chain engine around its launches
    int kpm_local = 1, kpm_rebase = 16;   // rebase: every so many sweeps the chain engine recomputes the trace sums from scratch
    const int32_t* kpm_f_cur = nullptr;
    const double* kpm_ks_in = nullptr;
    double* kpm_ks_out = nullptr;
    unsigned char* d_kpm_hop0 = nullptr;   // [N] hop distance from site 0
    double* d_ks_io = nullptr;             // [2][max_batch][FKMC_KPM_STATE] records of fkmc_logz_kpm_batched_local
    int32_t* d_f_ref = nullptr;            // [max_batch][N] its reference configurations
    bool kpm_state_written = false;        // the last KPM launch produced state records (two-kernel 2-D path)
    // fkmc_logz_kpm_batched_local: the reference configurations and records are uploaded on a second stream under the Lanczos kernel
    // (which only reads the proposals); the moments launch waits for kpm_wait_event when it is set
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_f_up = nullptr, ev_ref_up = nullptr, kpm_wait_event = nullptr;
    int* d_kpm_steps = nullptr;   // [max_batch] Lanczos steps of the last KPM launch (diagnostics)
    double* d_aux = nullptr;      // [max_batch][2][N] cached_exp / cached_fermi staging
    double* d_ev_scratch = nullptr;  // eigenvector path: tridiagonal eigenvectors | inverse-iteration factors | T factors (grown on demand)
    size_t ev_scratch_cap = 0;       // doubles

    // Chebyshev tables for the (M, G) last used
    int cheb_M = 0, cheb_G = 0;
    double* d_chebt = nullptr;    // [M][G]
    double* d_lobatto = nullptr;  // [G]
    double* d_dtheta = nullptr;   // [G-1]

    fkmc_chain_state chain;

    // end-of-run collective (gather.cu)
    void* nccl_comm = nullptr;  // ncclComm_t
    int nccl_nranks = 0, nccl_rank = 0;
    double* d_gather = nullptr;  // gathered + reordered series
    size_t gather_cap = 0;       // doubles per half

    // instrumentation
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool profiling = false;
    std::map<std::string, fkmc_profile_entry> prof;
    std::vector<fkmc_pending_event> pending;
};

// ---- error helpers ----
int fkmc_set_error(fkmc_ctx* ctx, int code, const std::string& msg);
#define FKMC_CUDA(ctx, call)                                                                              \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fkmc_set_error(ctx, FKMC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// RAII-less profiling scope: records events around a kernel family when profiling is on
struct fkmc_prof_scope {
    fkmc_ctx* ctx;
    const char* name;
    cudaEvent_t a = nullptr, b = nullptr;
    fkmc_prof_scope(fkmc_ctx* c, const char* n);
    ~fkmc_prof_scope();
};

// ---- workspaces ----
int fkmc_ensure_dense_ws(fkmc_ctx* ctx);

// ---- lattice (host) ----
int fkmc_build_lattice(fkmc_ctx* ctx);

// ---- kernel launchers (each returns an fkmc_status) ----
// dense H (lower triangle) from f: A[b] = hopping + diag(U f - mu_c)
int fkmc_launch_build_h(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double* d_A);
// blocked Householder tridiagonalisation, one CTA per matrix; A is overwritten
int fkmc_launch_sytrd(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_d, double* d_e, double* d_tau, double* d_W);
// two-stage tridiagonalisation
int fkmc_launch_sy2sb(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_AB);
// band path: lattice matrix in a folded ordering (half-bandwidth <= 64) -> half-bandwidth 8 (sb2sb.cu)
size_t fkmc_band_stride(int N);
int fkmc_band_setup(fkmc_ctx* ctx);
bool fkmc_use_band(const fkmc_ctx* ctx);
int fkmc_launch_band_reduce(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double* d_band, double* d_AB);
int fkmc_launch_sb2st(fkmc_ctx* ctx, const double* d_AB, int N, int B, double* d_d, double* d_e);
size_t fkmc_sy2sb_smem(int N);
size_t fkmc_sy2sb_scratch(int N);
size_t fkmc_sb2st_smem(int N);
int fkmc_launch_sy2sb_tiled(fkmc_ctx* ctx, double* d_At, int N, int B, double* d_AB);
int fkmc_launch_build_h_tiled(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double* d_At);
int fkmc_launch_to_tiled(fkmc_ctx* ctx, const double* d_A, int N, int B, double* d_At);
size_t fkmc_tiled_stride(int N);
bool fkmc_use_tiled(const fkmc_ctx* ctx, int N);
// f -> Hamiltonian -> tridiagonal (d, e) with the context's selected algorithm and matching matrix layout
int fkmc_build_tridiag(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double* d_d, double* d_e);
// dense -> tridiagonal with the context's selected algorithm (A is overwritten)
int fkmc_tridiagonalize(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_d, double* d_e);
// Sturm bisection + fused logZ / energy: out[b*8 + {0: logZ, 1: E_c, 2: d2E}]
int fkmc_launch_tridiag_eig(fkmc_ctx* ctx, const double* d_d, const double* d_e, int N, int B, double beta, double* d_evals,
                            long evals_stride, const int32_t* d_slot, long slot_stride, double* d_out, double* d_exp,
                            double* d_fermi);
int fkmc_launch_energy(fkmc_ctx* ctx, const double* d_evals, long evals_stride, const int32_t* d_slot, long slot_stride, int N,
                       int B, double beta, double* d_out);
// KPM: Lanczos extremal eigenvalues + Chebyshev moments + logZ
int fkmc_prepare_cheb(fkmc_ctx* ctx, int M, int G);
int fkmc_launch_kpm(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, int M, int G,
                    double* d_moments, double* d_ab, double* d_logz);
// two-kernel variant for the regular 2-D lattices (kpm2d.cu); slot_val = per-slot hopping constants
bool fkmc_kpm2d_applicable(const fkmc_ctx* ctx, int M);
int fkmc_kpm_prepare_local(fkmc_ctx* ctx);
int fkmc_launch_kpm2d(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, int M, int G, const double* slot_val,
                      double* d_moments, double* d_ab, double* d_logz);
// eigenvector path (measurement sweeps): evals/out on the device, eigenvectors and IPR to the host (or IPR to d_ipr)
int fkmc_eigvec_pipeline(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out,
                         double* h_evecs, double* h_ipr_host, double* d_ipr);
// eigenvector path with device outputs only: evals [B][N], out [B][8], vt [B][N][N] site-major (vt[b][i][k] = component i of eigenvector k)
int fkmc_eigvec_pipeline_dev2(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out, double* d_evecs,
                              double* d_vt, double* d_ipr);
int fkmc_eigvec_pipeline_dev(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out, double* d_vt);
// fast update of the dense moves (secular.cu)
int fkmc_fu_alloc(fkmc_ctx* ctx);
void fkmc_fu_free(fkmc_ctx* ctx);
int fkmc_fu_refresh(fkmc_ctx* ctx, int check);
int fkmc_fu_evaluate(fkmc_ctx* ctx);
int fkmc_fu_commit(fkmc_ctx* ctx);
int fkmc_fu_ipr(fkmc_ctx* ctx, double* d_ipr);
int fkmc_launch_secular_only(fkmc_ctx* ctx, int N, int B, const double* d_lam, const double* d_z, const double* d_rho, double* d_out);
// measure_stiffness on device configurations (stiffness.cu): d_st [B], d_cd [B][n_w]
int fkmc_stiffness_dev(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double offset, int n_w, const double* d_w,
                       double* d_st, double* d_cd);
// chains
int fkmc_chain_free(fkmc_ctx* ctx);
extern "C" int fkmc_comm_destroy(fkmc_ctx* ctx);
// reads and clears the non-convergence flag (synchronises the stream): FKMC_OK or FKMC_ERR_NOCONV
int fkmc_check_flag(fkmc_ctx* ctx);

#ifdef __CUDACC__
// ---- device helpers ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum; red must hold >= 33 doubles; all threads get the result.  Two barriers.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double s = lane < nw ? red[lane] : 0.0;
        s = warp_sum(s);
        if (lane == 0) red[32] = s;
    }
    __syncthreads();
    return red[32];
}
// FP64 tensor-core MMA (SASS: DMMA.8x8x4): D(8x8) += A(8x4) * B(4x8).
// lane = 4*g + t:  a = A[g][t], b = B[t][g], (c0, c1) = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
#endif
