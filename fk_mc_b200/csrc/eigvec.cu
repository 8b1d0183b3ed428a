// Eigenvector path, used on measurement sweeps only (IPR, eigenfunctions, stiffness):
// configuration_t::calc_ed(true) (src/configuration.cpp:213,216-219) and measure_ipr::accumulate
// (include/fk_mc/measures/ipr.hpp:39-56).
//
//   H --(one-stage blocked sytrd, reflectors kept in A)--> T --(Sturm bisection)--> eigenvalues
//     --(inverse iteration on T, tridiagonal LU with partial pivoting, one thread per eigenvalue,
//        modified Gram-Schmidt inside clusters of close eigenvalues)--> eigenvectors of T
//     --(back-transformation Z = H_0 H_1 ... H_{n-2} Z_T, 16 eigenvector columns per CTA held in shared
//        memory, IPR fused into the epilogue)--> eigenvectors of H, column k <-> eigenvalue k.
// The inverse iteration follows the structure of LAPACK dstein/dlagtf/dlagts (restated, not copied): it is
// what the reference's Eigen solver delivers up to the choice of basis inside degenerate subspaces.
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "common.cuh"

namespace {

constexpr int BT_COLS = 16;   // eigenvector columns per back-transformation CTA
constexpr int BT_THREADS = 256;

// scratch layout: element i of eigenvalue k of array q lives at ((q * N + i) * N + k)  (coalesced across k)
struct stein_args {
    const double* d;       // [B][N]
    const double* e;       // [B][N]
    const double* evals;   // [B][N] ascending
    double* zt;            // [B][N][N]  zt[i*N + k] = component i of tridiagonal eigenvector k
    double* scratch;       // [B][5][N][N]
    int N;
};

__device__ __forceinline__ double hash_unit(unsigned k, unsigned i, unsigned it) {
    unsigned h = k * 2654435761u ^ (i + 0x9e3779b9u) * 2246822519u ^ (it * 3266489917u);
    h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return (double)h * (2.0 / 4294967296.0) - 1.0;
}

// One thread per eigenvalue; the first eigenvalue of every cluster (gap to its predecessor > ortol) computes the whole
// cluster sequentially so that the Gram-Schmidt sweep can see the finished predecessors.
__global__ void __launch_bounds__(128) stein_kernel(stein_args P) {
    const int N = P.N, b = blockIdx.y;
    const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (k0 >= N) return;
    const double* d = P.d + (size_t)b * N;
    const double* e = P.e + (size_t)b * N;
    const double* w = P.evals + (size_t)b * N;
    double* zt = P.zt + (size_t)b * N * N;
    double* S = P.scratch + (size_t)b * 5 * N * N;
    const size_t NN = (size_t)N * N;
    double* fa = S + k0;            // pivots           fa[i*N]
    double* fb = S + NN + k0;       // first super-diagonal of U
    double* fc = S + 2 * NN + k0;   // multipliers
    double* fd = S + 3 * NN + k0;   // second super-diagonal of U
    double* fi = S + 4 * NN + k0;   // interchange flags (0 / 1)

    // 1-norm of T
    double onenrm = fabs(d[0]) + (N > 1 ? fabs(e[0]) : 0.0);
    for (int i = 1; i < N; ++i) onenrm = fmax(onenrm, fabs(d[i]) + fabs(e[i - 1]) + (i < N - 1 ? fabs(e[i]) : 0.0));
    if (onenrm == 0.0) onenrm = 1.0;
    const double eps = DBL_EPSILON;
    const double ortol = 1e-6 * onenrm;
    if (k0 > 0 && w[k0] - w[k0 - 1] <= ortol) return;  // not the first of its cluster

    double xjm = 0.0;
    for (int j = k0; j < N && (j == k0 || w[j] - w[j - 1] <= ortol); ++j) {
        double xj = w[j];
        if (j > k0) {  // keep the shifts of a cluster distinct (dstein: pertol = 10 eps |xj|)
            const double pertol = 10.0 * fabs(eps * xj) + 10.0 * eps * eps * onenrm;
            if (xj - xjm < pertol) xj = xjm + pertol;
        }
        xjm = xj;
        double* z = zt + j;  // z[i*N]
        // ---- LU factorisation of T - xj I with partial pivoting (row interchanges of adjacent rows) ----
        const double tol = eps * onenrm;
        double ak = d[0] - xj;                 // current pivot row: (ak, bk) on columns k, k+1
        double bk = (N > 1) ? e[0] : 0.0;
        for (int k = 0; k < N - 1; ++k) {
            const double ck = e[k];            // sub-diagonal entry of row k+1 in column k
            const double a1 = d[k + 1] - xj;   // row k+1: (ck, a1, b1)
            const double b1 = (k < N - 2) ? e[k + 1] : 0.0;
            if (fabs(ck) <= fabs(ak)) {        // no interchange
                const double mult = (ak != 0.0) ? ck / ak : 0.0;
                fa[(size_t)k * N] = ak;
                fb[(size_t)k * N] = bk;
                fd[(size_t)k * N] = 0.0;
                fc[(size_t)k * N] = mult;
                fi[(size_t)k * N] = 0.0;
                ak = a1 - mult * bk;
                bk = b1;
            } else {                           // interchange rows k and k+1
                const double mult = ak / ck;
                fa[(size_t)k * N] = ck;
                fb[(size_t)k * N] = a1;
                fd[(size_t)k * N] = b1;
                fc[(size_t)k * N] = mult;
                fi[(size_t)k * N] = 1.0;
                ak = bk - mult * a1;
                bk = -mult * b1;
            }
        }
        fa[(size_t)(N - 1) * N] = ak;
        fb[(size_t)(N - 1) * N] = 0.0;
        fd[(size_t)(N - 1) * N] = 0.0;
        // ---- inverse iteration ----
        for (int i = 0; i < N; ++i) z[(size_t)i * N] = hash_unit((unsigned)j, (unsigned)i, 0u);
        int good = 0;
        for (int its = 0; its < 6 && good < 2; ++its) {
            // scale the right-hand side (dstein: n * onenrm * max(eps, |last pivot|) / ||x||_1)
            double n1 = 0.0;
            for (int i = 0; i < N; ++i) n1 += fabs(z[(size_t)i * N]);
            const double scl = (double)N * onenrm * fmax(eps, fabs(fa[(size_t)(N - 1) * N])) / fmax(n1, DBL_MIN);
            for (int i = 0; i < N; ++i) z[(size_t)i * N] *= scl;
            // forward: apply the interchanges and multipliers
            for (int k = 0; k < N - 1; ++k) {
                const double yk = z[(size_t)k * N], yk1 = z[(size_t)(k + 1) * N], m = fc[(size_t)k * N];
                if (fi[(size_t)k * N] == 0.0) {
                    z[(size_t)(k + 1) * N] = yk1 - m * yk;
                } else {
                    z[(size_t)k * N] = yk1;
                    z[(size_t)(k + 1) * N] = yk - m * yk1;
                }
            }
            // backward: U x = y with tiny pivots replaced by +-tol
            double x1 = 0.0, x2 = 0.0;
            for (int k = N - 1; k >= 0; --k) {
                double piv = fa[(size_t)k * N];
                if (fabs(piv) < tol) piv = (piv < 0.0) ? -tol : tol;
                const double t = z[(size_t)k * N] - fb[(size_t)k * N] * x1 - fd[(size_t)k * N] * x2;
                const double x = t / piv;
                z[(size_t)k * N] = x;
                x2 = x1;
                x1 = x;
            }
            // re-orthogonalise against the finished vectors of this cluster (modified Gram-Schmidt)
            for (int q = k0; q < j; ++q) {
                const double* zq = zt + q;
                double dot = 0.0;
                for (int i = 0; i < N; ++i) dot = fma(z[(size_t)i * N], zq[(size_t)i * N], dot);
                for (int i = 0; i < N; ++i) z[(size_t)i * N] = fma(-dot, zq[(size_t)i * N], z[(size_t)i * N]);
            }
            // growth check + normalisation
            double nrm2 = 0.0, amax = 0.0;
            for (int i = 0; i < N; ++i) {
                const double x = z[(size_t)i * N];
                nrm2 = fma(x, x, nrm2);
                amax = fmax(amax, fabs(x));
            }
            if (!(nrm2 > 0.0) || !isfinite(nrm2)) {  // breakdown: restart from another pseudo-random vector
                for (int i = 0; i < N; ++i) z[(size_t)i * N] = hash_unit((unsigned)j, (unsigned)i, (unsigned)its + 1u);
                continue;
            }
            const double inv = 1.0 / sqrt(nrm2);
            for (int i = 0; i < N; ++i) z[(size_t)i * N] *= inv;
            if (amax >= sqrt(0.1 / (double)N)) ++good;  // the solve amplified the vector enough: count a converged step
        }
    }
}

// Z = H_0 ... H_{n-2} Z_T for a block of 16 eigenvector columns, then IPR.  A holds the reflectors of the one-stage
// sytrd (v_i in column i below the diagonal, leading 1 explicit at row i+1), tau[i] their scalars.
__global__ void __launch_bounds__(BT_THREADS, 1)
backtransform_kernel(const double* __restrict__ A_all, const double* __restrict__ tau_all, const double* __restrict__ zt_all, int N,
                     double* __restrict__ evecs_all /*[B][N][N] col-major or null*/, double* __restrict__ ipr_all /*[B][N] or null*/,
                     double* __restrict__ vt_all /*[B][N][N] site-major (vt[i][k]) or null*/) {
    extern __shared__ double zb[];  // [N][17] (padded rows)
    __shared__ double wpart[BT_THREADS / 32][BT_COLS];
    __shared__ double wfull[BT_COLS];
    const int LDZ = BT_COLS + 1;
    const int b = blockIdx.y, c0 = blockIdx.x * BT_COLS, tid = threadIdx.x;
    const int c = tid & (BT_COLS - 1), rl = tid >> 4;  // column inside the block, row lane (0..15)
    const int lane = tid & 31, warp = tid >> 5;
    const double* A = A_all + (size_t)b * N * N;
    const double* tau = tau_all + (size_t)b * N;
    const double* zt = zt_all + (size_t)b * N * N;
    const int ncol = min(BT_COLS, N - c0);
    for (int idx = tid; idx < N * BT_COLS; idx += BT_THREADS) {
        const int i = idx / BT_COLS, cc = idx % BT_COLS;
        zb[i * LDZ + cc] = (cc < ncol) ? zt[(size_t)i * N + c0 + cc] : 0.0;
    }
    __syncthreads();
    for (int i = N - 2; i >= 0; --i) {
        const double t = tau[i];
        if (t == 0.0) continue;  // uniform
        const double* v = A + (size_t)i * N;  // v[r], r > i; v[i+1] = 1
        // row ownership must not depend on i: thread (c, rl) owns the rows r == rl (mod 16) of column c for the whole kernel
        const int rs = i + 1 + ((rl - (i + 1)) & 15);
        double part = 0.0;
        for (int r = rs; r < N; r += 16) part = fma(v[r], zb[r * LDZ + c], part);
        part += __shfl_xor_sync(0xffffffffu, part, 16);  // the two row lanes of this warp
        if (lane < 16) wpart[warp][c] = part;
        __syncthreads();
        if (tid < BT_COLS) {
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < BT_THREADS / 32; ++ww) s += wpart[ww][tid];
            wfull[tid] = t * s;
        }
        __syncthreads();
        const double wc = wfull[c];
        for (int r = rs; r < N; r += 16) zb[r * LDZ + c] = fma(-v[r], wc, zb[r * LDZ + c]);
        // (element (r, c) is only ever touched by thread (c, r mod 16), so no barrier is needed before the next reflector)
    }
    __syncthreads();
    if (evecs_all) {
        double* ev = evecs_all + (size_t)b * N * N;
        for (int idx = tid; idx < N * BT_COLS; idx += BT_THREADS) {
            const int cc = idx / N, i = idx % N;
            if (cc < ncol) ev[(size_t)(c0 + cc) * N + i] = zb[i * LDZ + cc];
        }
    }
    if (vt_all) {
        double* vt = vt_all + (size_t)b * N * N;
        for (int idx = tid; idx < N * BT_COLS; idx += BT_THREADS) {
            const int i = idx / BT_COLS, cc = idx % BT_COLS;
            if (cc < ncol) vt[(size_t)i * N + c0 + cc] = zb[i * LDZ + cc];
        }
    }
    if (ipr_all) {
        // ipr_k = ||psi||_4 / ||psi||_2^2  (include/fk_mc/measures/ipr.hpp:47-53)
        double s2 = 0.0, s4 = 0.0;
        for (int r = rl; r < N; r += 16) {
            const double x = zb[r * LDZ + c], x2 = x * x;
            s2 += x2;
            s4 = fma(x2, x2, s4);
        }
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
        s4 += __shfl_xor_sync(0xffffffffu, s4, 16);
        __shared__ double p2[BT_THREADS / 32][BT_COLS], p4[BT_THREADS / 32][BT_COLS];
        if (lane < 16) { p2[warp][c] = s2; p4[warp][c] = s4; }
        __syncthreads();
        if (tid < ncol) {
            double a2 = 0.0, a4 = 0.0;
            for (int ww = 0; ww < BT_THREADS / 32; ++ww) { a2 += p2[ww][tid]; a4 += p4[ww][tid]; }
            ipr_all[(size_t)b * N + c0 + tid] = sqrt(sqrt(a4)) / a2;
        }
    }
}

}  // namespace

// Eigen-decomposition of the Hamiltonians of B configurations (device f).  Outputs on the device: evals [B][N],
// optionally evecs [B][N][N] (column-major) and ipr [B][N]; logZ etc. in d_out [B][8].  Processes the batch in chunks.
static int eigvec_pipeline_impl(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out,
                                double* h_evecs, double* h_ipr_host, double* d_ipr, double* d_vt) {
    const int N = ctx->N;
    const size_t NN = (size_t)N * N;
    if (sizeof(double) * (size_t)N * (BT_COLS + 1) + 4096 > ctx->smem_optin)
        return fkmc_set_error(ctx, FKMC_ERR_INVALID, "eigenvectors: N too large for the back-transformation kernel");
    // chunk so that scratch (6 N^2 doubles per matrix) stays below ~6 GB
    int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, (size_t)(6.0e9 / (6.0 * NN * 8.0))));
    chunk = std::min(chunk, ctx->max_batch);
    double *d_zt = nullptr, *d_scr = nullptr, *d_ev = nullptr, *d_ip = nullptr;
    FKMC_CUDA(ctx, cudaMalloc(&d_zt, sizeof(double) * NN * chunk));
    FKMC_CUDA(ctx, cudaMalloc(&d_scr, sizeof(double) * 5 * NN * chunk));
    if (h_evecs) FKMC_CUDA(ctx, cudaMalloc(&d_ev, sizeof(double) * NN * chunk));
    if (!d_ipr && h_ipr_host) FKMC_CUDA(ctx, cudaMalloc(&d_ip, sizeof(double) * (size_t)N * chunk));
    int rc = FKMC_OK;
    const size_t smem = sizeof(double) * (size_t)N * (BT_COLS + 1);
    cudaFuncSetAttribute(backtransform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int b0 = 0; b0 < B && !rc; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        const int32_t* f = d_f + (size_t)b0 * N;
        // one-stage tridiagonalisation keeps the reflectors needed for the back-transformation
        if ((rc = fkmc_launch_build_h(ctx, f, nb, U, mu_c, ctx->d_A))) break;
        if ((rc = fkmc_launch_sytrd(ctx, ctx->d_A, N, nb, ctx->d_d, ctx->d_e, ctx->d_tau, ctx->d_W))) break;
        if ((rc = fkmc_launch_tridiag_eig(ctx, ctx->d_d, ctx->d_e, N, nb, beta, d_evals + (size_t)b0 * N, N, nullptr, 0, d_out + (size_t)b0 * 8,
                                          nullptr, nullptr)))
            break;
        {
            fkmc_prof_scope ps(ctx, "stein");
            stein_args P{ctx->d_d, ctx->d_e, d_evals + (size_t)b0 * N, d_zt, d_scr, N};
            dim3 grid((N + 127) / 128, nb);
            stein_kernel<<<grid, 128, 0, ctx->stream>>>(P);
            ctx->launches++;
        }
        double* iprp = d_ipr ? d_ipr + (size_t)b0 * N : d_ip;
        {
            fkmc_prof_scope ps(ctx, "backtransform");
            dim3 grid((N + BT_COLS - 1) / BT_COLS, nb);
            backtransform_kernel<<<grid, BT_THREADS, smem, ctx->stream>>>(ctx->d_A, ctx->d_tau, d_zt, N, d_ev, iprp, d_vt ? d_vt + (size_t)b0 * NN : nullptr);
            ctx->launches++;
        }
        if (cudaGetLastError() != cudaSuccess) { rc = fkmc_set_error(ctx, FKMC_ERR_CUDA, "eigenvector kernels failed to launch"); break; }
        if (h_evecs) cudaMemcpyAsync(h_evecs + (size_t)b0 * NN, d_ev, sizeof(double) * NN * nb, cudaMemcpyDeviceToHost, ctx->stream);
        if (h_ipr_host && !d_ipr) cudaMemcpyAsync(h_ipr_host + (size_t)b0 * N, d_ip, sizeof(double) * (size_t)N * nb, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = fkmc_set_error(ctx, FKMC_ERR_CUDA, cudaGetErrorString(cudaGetLastError())); break; }
    }
    cudaFree(d_zt); cudaFree(d_scr); cudaFree(d_ev); cudaFree(d_ip);
    return rc;
}

int fkmc_eigvec_pipeline(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out,
                         double* h_evecs, double* h_ipr_host, double* d_ipr) {
    return eigvec_pipeline_impl(ctx, d_f, B, U, mu_c, beta, d_evals, d_out, h_evecs, h_ipr_host, d_ipr, nullptr);
}

int fkmc_eigvec_pipeline_dev(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out, double* d_vt) {
    int rc = fkmc_ensure_dense_ws(ctx);
    if (rc) return rc;
    return eigvec_pipeline_impl(ctx, d_f, B, U, mu_c, beta, d_evals, d_out, nullptr, nullptr, nullptr, d_vt);
}
