// measure_stiffness::accumulate (include/fk_mc/measures/stiffness.hpp:129-187) on the GPU: Drude weight and optical conductivity
// from the eigen-decomposition of a configuration's Hamiltonian (hypercubic lattices, D >= 2; current along the first coordinate).
//
//   Tm(i, i -+ x) = -t,  Jm(i, i - x) = -1, Jm(i, i + x) = +1                                    (stiffness.hpp:84-126)
//   T  = -pi sum_k (V^T Tm V)_kk f_k
//   mJ = V^T Jm V                                                                              (the N^3 contraction: DMMA GEMM)
//   V  = sum_{i > j, |e_i - e_j| > 1e-12, |sigma_ij| > 1e-13} 2 sigma_ij / (e_i - e_j),   sigma_ij = pi (f_j - f_i) mJ(j,i) mJ(i,j)
//   stiffness = (V + T) / N;   cond(w) = sum over the same pairs of L(w; e_j - e_i, sigma) + L(w; e_i - e_j, -sigma),
//   L(w; x0, s) = offset / pi / ((w + x0)^2 + offset^2) s                                     (resonant_term, stiffness.hpp:15-29)
// The eigenvectors never leave the device: eigvec.cu delivers them in both layouts (eigenvector-major for the left factor of the
// GEMM, site-major for the stencil Jm V), gemm_tile.cuh does V^T (Jm V), one CTA per matrix does the O(N^2) Kubo sums.
#include <algorithm>

#include "common.cuh"
#include "gemm_tile.cuh"

namespace {

// JV[i][k] = -vt[left(i)][k] + vt[right(i)][k];  tdiag partial: sum_i vt[i][k] (-t)(vt[left][k] + vt[right][k])
// grid (ceil(N / 256) column chunks, B); thread = eigenvector k, loop over sites i (coalesced along k)
__global__ void __launch_bounds__(256) jv_kernel(const double* __restrict__ vt_all, int N, int stride_x, int Lx, double t_hop, double* __restrict__ jv_all,
                                                 double* __restrict__ tdiag_all) {
    const int b = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
    if (k >= N) return;
    const double* vt = vt_all + (size_t)b * N * N;
    double* jv = jv_all + (size_t)b * N * N;
    double td = 0.0;
    for (int i = 0; i < N; ++i) {
        // first coordinate x = i / stride_x (index = sum pos[d] prod_{j>d} L: the FIRST coordinate is the slowest, hypercubic.cpp:31-51)
        const int x = i / stride_x;
        const int il = i + (x == 0 ? (Lx - 1) * stride_x : -stride_x);
        const int ir = i + (x == Lx - 1 ? -(Lx - 1) * stride_x : stride_x);
        const double vl = vt[(size_t)il * N + k], vr = vt[(size_t)ir * N + k], v0 = vt[(size_t)i * N + k];
        jv[(size_t)i * N + k] = vr - vl;
        td = fma(v0, -t_hop * (vl + vr), td);
    }
    tdiag_all[(size_t)b * N + k] = td;
}

__global__ void __launch_bounds__(256, 2) gemm_nn_kernel(const double* __restrict__ A_all, const double* __restrict__ B_all, double* __restrict__ C_all, int N) {
    extern __shared__ __align__(16) double sm[];
    const size_t NN = (size_t)N * N;
    fkgemm::tile(A_all + blockIdx.y * NN, B_all + blockIdx.y * NN, C_all + blockIdx.y * NN, N, blockIdx.x, sm);
}

// Kubo sums of one matrix per CTA.  out: stiffness[b], cond[b][n_w]
__global__ void __launch_bounds__(1024) kubo_kernel(const double* __restrict__ mj_all, const double* __restrict__ evals_all, const double* __restrict__ tdiag_all, int N,
                                                    double beta, double offset, int n_w, const double* __restrict__ wgrid, double* __restrict__ stiff,
                                                    double* __restrict__ cond) {
    extern __shared__ double ksm[];   // ev[N] fermi[N] red[40]
    double* ev = ksm;
    double* fe = ev + N;
    double* red = fe + N;
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const double* mJ = mj_all + (size_t)b * N * N;
    for (int k = tid; k < N; k += T) {
        const double e = evals_all[(size_t)b * N + k];
        ev[k] = e;
        fe[k] = 1.0 / (1.0 + exp(beta * e));   // ed_cache::cached_fermi (configuration.cpp:237)
    }
    __syncthreads();
    double Tsum = 0.0;
    for (int k = tid; k < N; k += T) Tsum = fma(tdiag_all[(size_t)b * N + k], fe[k], Tsum);
    double Vsum = 0.0;
    double cw[FKMC_MAX_W];  // up to 8 frequencies accumulated per pass
    for (int w0 = 0; w0 < max(n_w, 1); w0 += FKMC_MAX_W) {
        const int nw = min(FKMC_MAX_W, n_w - w0);
#pragma unroll
        for (int w = 0; w < FKMC_MAX_W; ++w) cw[w] = 0.0;
        // pairs i > j: row i handled by threads along j (mJ(i, j) coalesced; mJ(j, i) strided but L2-resident)
        for (int i = 1; i < N; ++i) {
            const double ei = ev[i], fi = fe[i];
            for (int j = tid; j < i; j += T) {
                const double de = ei - ev[j];
                if (!(fabs(de) > 1e-12)) continue;
                const double sigma = M_PI * (fe[j] - fi) * mJ[(size_t)j * N + i] * mJ[(size_t)i * N + j];
                if (!(fabs(sigma) > 1e-13)) continue;
                if (w0 == 0) Vsum += 2.0 * sigma / de;
                for (int w = 0; w < nw; ++w) {
                    const double x1 = wgrid[w0 + w] - de, x2 = wgrid[w0 + w] + de;   // resonant_term(e_j - e_i, sigma) and (e_i - e_j, -sigma): lorentzian(w + energy)
                    cw[w] += offset / M_PI * sigma * (1.0 / (x1 * x1 + offset * offset) - 1.0 / (x2 * x2 + offset * offset));
                }
            }
        }
        for (int w = 0; w < nw; ++w) {
            const double s = block_sum(cw[w], red);
            if (tid == 0) cond[(size_t)b * n_w + w0 + w] = s;
            __syncthreads();
        }
    }
    const double Tt = block_sum(Tsum, red);
    __syncthreads();
    const double Vt = block_sum(Vsum, red);
    if (tid == 0) stiff[b] = (Vt - M_PI * Tt) / (double)N;
}

}  // namespace

// Device-side core: configurations d_f [B][N] (device) -> d_st [B], d_cd [B][n_w] (device).  d_w: the frequency grid on the device.
int fkmc_stiffness_dev(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double offset, int n_w, const double* d_w,
                       double* d_st, double* d_cd) {
    if (ctx->kind != FKMC_CUBIC2D && ctx->kind != FKMC_CUBIC3D)
        return fkmc_set_error(ctx, FKMC_ERR_INVALID, "stiffness: hypercubic lattices with D >= 2 only (stiffness.hpp:147)");
    const int N = ctx->N, L = ctx->L;
    const size_t NN = (size_t)N * N;
    const int stride_x = N / L;  // L^(D-1): the first coordinate is the slowest
    // chunk: 4 N^2 doubles per matrix here (+ the eigenvector pipeline's own scratch)
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)B, (size_t)(4.0e9 / (4.0 * NN * 8.0))));
    double *d_ev = nullptr, *d_vt = nullptr, *d_jv = nullptr, *d_mj = nullptr, *d_td = nullptr, *d_evals = nullptr;
    auto cleanup = [&]() { cudaFree(d_ev); cudaFree(d_vt); cudaFree(d_jv); cudaFree(d_mj); cudaFree(d_td); cudaFree(d_evals); };
#define FKMC_ST(call)                                                                                       \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) { cleanup(); return fkmc_set_error(ctx, FKMC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } \
    } while (0)
    FKMC_ST(cudaMalloc(&d_ev, sizeof(double) * NN * chunk));
    FKMC_ST(cudaMalloc(&d_vt, sizeof(double) * NN * chunk));
    FKMC_ST(cudaMalloc(&d_jv, sizeof(double) * NN * chunk));
    FKMC_ST(cudaMalloc(&d_mj, sizeof(double) * NN * chunk));
    FKMC_ST(cudaMalloc(&d_td, sizeof(double) * (size_t)N * chunk));
    FKMC_ST(cudaMalloc(&d_evals, sizeof(double) * (size_t)N * chunk));
    FKMC_ST(cudaFuncSetAttribute(gemm_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fkgemm::smem_bytes()));
    int rc = FKMC_OK;
    for (int b0 = 0; b0 < B && !rc; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        if ((rc = fkmc_eigvec_pipeline_dev2(ctx, d_f + (size_t)b0 * N, nb, U, mu_c, beta, d_evals, ctx->d_out, d_ev, d_vt, nullptr))) break;
        {
            fkmc_prof_scope ps(ctx, "stiffness_jv");
            jv_kernel<<<dim3((N + 255) / 256, nb), 256, 0, ctx->stream>>>(d_vt, N, stride_x, L, ctx->t, d_jv, d_td);
            ctx->launches++;
        }
        {
            fkmc_prof_scope ps(ctx, "stiffness_gemm");   // mJ = V^T (Jm V): A = eigenvector-major evecs [k][i], B = JV [i][l]
            gemm_nn_kernel<<<dim3(fkgemm::tiles(N), nb), 256, fkgemm::smem_bytes(), ctx->stream>>>(d_ev, d_jv, d_mj, N);
            ctx->launches++;
        }
        {
            fkmc_prof_scope ps(ctx, "stiffness_kubo");
            const int T = std::min(1024, ((N + 31) / 32) * 32);
            kubo_kernel<<<nb, T, sizeof(double) * (2 * (size_t)N + 40), ctx->stream>>>(d_mj, d_evals, d_td, N, beta, offset, n_w, d_w, d_st + b0,
                                                                                       d_cd + (size_t)b0 * n_w);
            ctx->launches++;
        }
        FKMC_ST(cudaGetLastError());
        FKMC_ST(cudaStreamSynchronize(ctx->stream));   // the scratch of this chunk is reused by the next one
    }
#undef FKMC_ST
    cleanup();
    return rc;
}

extern "C" int fkmc_stiffness_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta, double offset, int n_w,
                                      const double* wgrid, double* stiffness, double* cond) {
    if (!ctx || !f || !stiffness || n_w < 0 || (n_w > 0 && (!wgrid || !cond))) return FKMC_ERR_INVALID;
    if (B < 1 || B > ctx->max_batch) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "B must be in [1, max_batch]");
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    const int N = ctx->N;
    double *d_w = nullptr, *d_st = nullptr, *d_cd = nullptr;
    auto cleanup = [&]() { cudaFree(d_w); cudaFree(d_st); cudaFree(d_cd); };
    if (cudaMalloc(&d_w, sizeof(double) * std::max(n_w, 1)) != cudaSuccess || cudaMalloc(&d_st, sizeof(double) * B) != cudaSuccess ||
        cudaMalloc(&d_cd, sizeof(double) * (size_t)B * std::max(n_w, 1)) != cudaSuccess) {
        cleanup();
        return fkmc_set_error(ctx, FKMC_ERR_CUDA, "stiffness: out of device memory");
    }
    cudaMemcpyAsync(ctx->d_f, f, sizeof(int32_t) * (size_t)B * N, cudaMemcpyHostToDevice, ctx->stream);
    if (n_w) cudaMemcpyAsync(d_w, wgrid, sizeof(double) * n_w, cudaMemcpyHostToDevice, ctx->stream);
    int rc = fkmc_stiffness_dev(ctx, ctx->d_f, B, U, mu_c, beta, offset, n_w, d_w, d_st, d_cd);
    if (!rc) {
        cudaMemcpyAsync(stiffness, d_st, sizeof(double) * B, cudaMemcpyDeviceToHost, ctx->stream);
        if (n_w) cudaMemcpyAsync(cond, d_cd, sizeof(double) * (size_t)B * n_w, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fkmc_set_error(ctx, FKMC_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
    }
    cleanup();
    if (rc) return rc;
    return fkmc_check_flag(ctx);
}
