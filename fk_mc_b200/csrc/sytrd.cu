// Batched blocked Householder tridiagonalisation (eigenvalues-only path), one CTA per matrix.
//
// Replaces the tridiagonalisation stage of Eigen::SelfAdjointEigenSolver as called from
// configuration_t::calc_ed (src/configuration.cpp:212-213).  The reference runs it unblocked
// (SYMV + SYR2 per column); here columns are processed in panels of NB = 32 (LAPACK dlatrd
// shape): inside a panel every column needs one SYMV with the trailing matrix as stored plus
// rank-2j corrections from the panel's (V, W); after the panel the trailing matrix receives the
// rank-2NB update  A22 -= V W^T + W V^T  on the FP64 tensor cores (mma.sync m8n8k4 -> DMMA.8x8x4).
//
// Layout: A column-major, lda = N, lower triangle live.  Reflector q of a panel is stored in column
// k0+q below the diagonal with its leading 1 written explicitly; W is an N x NB column-major
// scratch panel per matrix.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int NB = FKMC_SYTRD_NB;

struct sytrd_smem {
    double* xs;   // [Np] current Householder vector (0 above row i+1)
    double* y1;   // [Np] SYMV: contributions to the task's own tile
    double* y2;   // [Np] SYMV: contributions to the partner tile
    double* col;  // [Np] updated column
    double* red;  // [40]
    double* tV;   // [NB]
    double* tW;   // [NB]
    double* rowV; // [NB]
    double* rowW; // [NB]
};

// transposed warp reduction: every lane holds vals[0..31]; on return lane c holds sum over lanes of vals[c].
__device__ __forceinline__ double warp_transpose_sum(double (&vals)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
            const double send = upper ? vals[k] : vals[k + off];
            const double keep = upper ? vals[k + off] : vals[k];
            vals[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return vals[0];
}

// One 32x32 tile task of the lower-triangle SYMV.  (a = own tile, p = partner tile.)
__device__ __forceinline__ void symv_tile(const double* __restrict__ A, int lda, int N, int a, int p, bool diag, int lane,
                                          const double* __restrict__ xs, double* y1, double* y2) {
    const int Rmax = a > p ? a : p, Cmin = a > p ? p : a;
    const int r = 32 * Rmax + lane;
    const bool rok = r < N;
    const double xr = xs[r];
    double vals[32];
    const double* Ap = A + (size_t)(32 * Cmin) * lda + r;
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) {
        const bool ok = rok && (32 * Cmin + cc < N) && (!diag || lane >= cc);
        vals[cc] = ok ? Ap[(size_t)cc * lda] : 0.0;
    }
    double rowacc = 0.0;
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) {
        rowacc = fma(vals[cc], xs[32 * Cmin + cc], rowacc);
        vals[cc] = (diag && lane == cc) ? 0.0 : vals[cc] * xr;
    }
    const double colsum = warp_transpose_sum(vals, lane);
    if (diag) {
        y1[32 * a + lane] = rowacc + colsum;
    } else if (a == Rmax) {
        y1[32 * a + lane] += rowacc;
        y2[32 * p + lane] += colsum;
    } else {
        y1[32 * a + lane] += colsum;
        y2[32 * p + lane] += rowacc;
    }
}

__global__ void __launch_bounds__(512, 1)
sytrd_lower_kernel(double* __restrict__ A_all, int N, double* __restrict__ d_all, double* __restrict__ e_all,
                   double* __restrict__ tau_all, double* __restrict__ W_all) {
    extern __shared__ double smem[];
    const int NT = (N + 31) >> 5, Np = NT * 32;
    sytrd_smem S;
    S.xs = smem;
    S.y1 = S.xs + Np;
    S.y2 = S.y1 + Np;
    S.col = S.y2 + Np;
    S.red = S.col + Np;
    S.tV = S.red + 40;
    S.tW = S.tV + NB;
    S.rowV = S.tW + NB;
    S.rowW = S.rowV + NB;

    const int b = blockIdx.x;
    const int lda = N, ldw = N;
    double* A = A_all + (size_t)b * N * N;
    double* W = W_all + (size_t)b * N * NB;
    double* dd = d_all + (size_t)b * N;
    double* ee = e_all + (size_t)b * N;
    double* tt = tau_all + (size_t)b * N;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;

    for (int r = tid; r < Np; r += T) { S.xs[r] = 0.0; S.y1[r] = 0.0; S.y2[r] = 0.0; S.col[r] = 0.0; }
    __syncthreads();

    for (int k0 = 0; k0 < N - 1; k0 += NB) {
        const int nbp = min(NB, N - 1 - k0);
        for (int j = 0; j < nbp; ++j) {
            const int i = k0 + j;
            // ---- (1) bring column i up to date with the panel's previous reflectors ----
            if (tid < j) {
                S.rowW[tid] = W[i + (size_t)tid * ldw];
                S.rowV[tid] = A[i + (size_t)(k0 + tid) * lda];
            }
            __syncthreads();
            for (int r = i + tid; r < N; r += T) {
                double a = A[r + (size_t)i * lda];
                for (int q = 0; q < j; ++q)
                    a -= A[r + (size_t)(k0 + q) * lda] * S.rowW[q] + W[r + (size_t)q * ldw] * S.rowV[q];
                S.col[r] = a;
            }
            __syncthreads();
            // ---- (2) Householder reflector annihilating col[i+2:] ----
            double part = 0.0;
            for (int r = i + 2 + tid; r < N; r += T) part = fma(S.col[r], S.col[r], part);
            const double tail2 = block_sum(part, S.red);
            const double c0 = S.col[i + 1];
            double tau, beta, inv;
            if (tail2 <= DBL_MIN) {
                tau = 0.0; beta = c0; inv = 0.0;
            } else {
                beta = sqrt(fma(c0, c0, tail2));
                if (c0 >= 0.0) beta = -beta;
                inv = 1.0 / (c0 - beta);
                tau = (beta - c0) / beta;
            }
            const int T0 = (i + 1) >> 5;
            for (int r = 32 * T0 + tid; r < N; r += T) {
                double v = 0.0;
                if (r == i + 1) v = 1.0;
                else if (r > i + 1) v = S.col[r] * inv;
                S.xs[r] = v;
                if (r > i) A[r + (size_t)i * lda] = v;
                S.y2[r] = 0.0;
            }
            if (tid == 0) {
                dd[i] = S.col[i];
                ee[i] = beta;
                tt[i] = tau;
            }
            __syncthreads();
            // ---- (3b) tV = V^T v, tW = W^T v over rows > i (one warp per dot product) ----
            for (int id = warp; id < 2 * j; id += nwarps) {
                const int q = id < j ? id : id - j;
                const double* src = id < j ? (A + (size_t)(k0 + q) * lda) : (W + (size_t)q * ldw);
                double s = 0.0;
                for (int r = i + 1 + lane; r < N; r += 32) s = fma(src[r], S.xs[r], s);
                s = warp_sum(s);
                if (lane == 0) { if (id < j) S.tV[q] = s; else S.tW[q] = s; }
            }
            // ---- (3a) y = A22 v with the trailing matrix as stored (lower triangle, 32x32 warp tiles) ----
            const int nt = NT - T0;
            for (int wt = warp; wt < nt; wt += nwarps) symv_tile(A, lda, N, T0 + wt, T0 + wt, true, lane, S.xs, S.y1, S.y2);
            const int smax = nt >> 1;
            for (int s = 1; s <= smax; ++s) {
                __syncthreads();
                const int lim = (2 * s == nt) ? (nt >> 1) : nt;
                for (int wt = warp; wt < lim; wt += nwarps) {
                    int pt = wt - s;
                    if (pt < 0) pt += nt;
                    symv_tile(A, lda, N, T0 + wt, T0 + pt, false, lane, S.xs, S.y1, S.y2);
                }
            }
            __syncthreads();
            // ---- (3c) + (4) w = tau (y - V tW - W tV);  w += -tau/2 (w.v) v ----
            part = 0.0;
            for (int r = i + 1 + tid; r < N; r += T) {
                double y = S.y1[r] + S.y2[r];
                for (int q = 0; q < j; ++q)
                    y -= A[r + (size_t)(k0 + q) * lda] * S.tW[q] + W[r + (size_t)q * ldw] * S.tV[q];
                y *= tau;
                S.col[r] = y;
                part = fma(y, S.xs[r], part);
            }
            const double wv = block_sum(part, S.red);
            const double alpha = -0.5 * tau * wv;
            for (int r = i + 1 + tid; r < N; r += T) W[r + (size_t)j * ldw] = fma(alpha, S.xs[r], S.col[r]);
            __syncthreads();
        }
        // ---- trailing update on FP64 tensor cores: A[r,c] -= sum_q V[r,q] W[c,q] + W[r,q] V[c,q], r >= c >= k1 ----
        const int k1 = k0 + nbp;
        if (k1 < N) {
            const int Tk = k1 >> 5, ntk = NT - Tk, ntl = ntk * (ntk + 1) / 2;
            const int g = lane >> 2, tg = lane & 3;
            for (int t = warp; t < ntl; t += nwarps) {
                int R = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
                while (R * (R + 1) / 2 > t) --R;
                while ((R + 1) * (R + 2) / 2 <= t) ++R;
                const int C = t - R * (R + 1) / 2;
                const int r0 = 32 * (Tk + R), c0 = 32 * (Tk + C);
                double acc[4][4][2];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
#pragma unroll 2
                for (int ks = 0; ks < (2 * NB) / 4; ++ks) {
                    const int kk = 4 * ks + tg;
                    const bool first = kk < NB;
                    const int q = first ? kk : kk - NB;
                    const bool qok = q < nbp;
                    // P = [V W] rows (A operand), Q = [W V] rows (B operand)
                    const double* Pcol = first ? (A + (size_t)(k0 + q) * lda) : (W + (size_t)q * ldw);
                    const double* Qcol = first ? (W + (size_t)q * ldw) : (A + (size_t)(k0 + q) * lda);
                    double af[4], bf[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int rr = r0 + 8 * x + g, cc = c0 + 8 * x + g;
                        af[x] = (qok && rr < N) ? Pcol[rr] : 0.0;
                        bf[x] = (qok && cc < N) ? Qcol[cc] : 0.0;
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int rr = r0 + 8 * x + g, cc = c0 + 8 * y + 2 * tg + h;
                            if (rr < N && cc < N && rr >= cc && cc >= k1) A[rr + (size_t)cc * lda] -= acc[x][y][h];
                        }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        dd[N - 1] = A[(size_t)(N - 1) * lda + (N - 1)];
        ee[N - 1] = 0.0;
        tt[N - 1] = 0.0;
    }
}

}  // namespace

int fkmc_launch_sytrd(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_d, double* d_e, double* d_tau, double* d_W) {
    fkmc_prof_scope ps(ctx, "sytrd");
    const int NT = (N + 31) / 32;
    int nwarps = NT < 16 ? NT : 16;
    if (nwarps < 2) nwarps = 2;
    const size_t smem = sizeof(double) * (4 * (size_t)NT * 32 + 40 + 4 * NB);
    if (smem > 48 * 1024) FKMC_CUDA(ctx, cudaFuncSetAttribute(sytrd_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sytrd_lower_kernel<<<B, nwarps * 32, smem, ctx->stream>>>(d_A, N, d_d, d_e, d_tau, d_W);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}
