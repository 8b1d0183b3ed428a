// (Small-matrix variant, N < 512: direct global fragment loads, 16 warps, no staging buffers.)
// Stage 1 of the two-stage tridiagonalisation: dense symmetric (lower) -> symmetric band with
// half-bandwidth 8, one CTA per matrix, every O(N^3) flop on the FP64 tensor cores.
//
// Together with sb2st.cu this replaces the tridiagonalisation stage of Eigen::SelfAdjointEigenSolver
// as called from configuration_t::calc_ed (src/configuration.cpp:212-213).  For block column k
// (columns k0..k0+7, trailing rows r0 = k0+8 .. N-1, m = N - r0):
//   1. Householder QR of the m x 8 panel in shared memory  ->  V (unit lower trapezoidal), tau, R
//   2. T (8x8, compact WY: Q = I - V T V^T) from the Gram matrix V^T V (DMMA)
//   3. Y0 = A22 V          -- SYMM over the stored lower triangle, 32x32 warp tiles, DMMA m8n8k4;
//                             tile tasks are paired cyclically so that no two warps add into the
//                             same rows of Y in the same step (no atomics)
//   4. Y = Y0 T,  X = V^T Y (DMMA),  Z = Y - 1/2 V (T^T X)
//   5. A22 -= V Z^T + Z V^T  -- rank-16 SYR2K on the lower triangle, DMMA, operands from shared memory
// R and the diagonal blocks are emitted in band storage AB[d][c] = A(c+d, c), d = 0..8.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int NB = 8;

struct s1_smem {
    double* V;    // [8][ld] column-major panel / reflectors
    double* Y;    // [8][ld] Y0 (own contributions) -> Y -> Z
    double* Y2;   // [8][ld] Y0 (partner contributions)
    double* G;    // [64] Gram / X / scratch
    double* Tm;   // [64] T
    double* M2;   // [64] T^T X
    double* Rs;   // [64] R
    double* tau;  // [8]
    double* red;  // [16*64 + 72]
    int ld;
};

// block-wide sum of K values per thread; result broadcast to all threads via out[0..K)
template <int K>
__device__ __forceinline__ void block_sum_vec(double (&v)[K], double* red, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double s = warp_sum(v[k]);
        if (lane == 0) red[warp * K + k] = s;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[w * K + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// C(8x8) = P^T Q over rows [0, m): P, Q column-major [8][ld] in shared memory.  Each warp accumulates a
// slice of rows with DMMA, partial tiles are summed through red[nwarps][64]; result in out[a*8 + c].
__device__ __forceinline__ void gram8(const double* P, const double* Q, int ld, int m, double* red, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    for (int r = 4 * warp; r < m; r += 4 * nw) {
        // A(g, k=t) = P[r+t][g], B(k=t, n=g) = Q[r+t][g]   (rows >= m are zero-padded)
        dmma884(c0, c1, P[g * ld + r + t], Q[g * ld + r + t]);
    }
    red[warp * 64 + g * 8 + 2 * t] = c0;
    red[warp * 64 + g * 8 + 2 * t + 1] = c1;
    __syncthreads();
    if (threadIdx.x < 64) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[w * 64 + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// One SYMM tile task: stored tile (Rmax, Cmin) of the trailing matrix (tile coordinates relative to r0).
// Contribution 1: Y[Rmax rows] += S V[Cmin rows];  contribution 2: Y[Cmin rows] += S^T V[Rmax rows].
__device__ __forceinline__ void symm_tile(const double* __restrict__ A22, int lda, int m, int a, int p, const s1_smem& S, int lane) {
    const int g = lane >> 2, t = lane & 3, ld = S.ld;
    const bool diag = (a == p);
    const int Rmax = a > p ? a : p, Cmin = a > p ? p : a;
    const int rb0 = 32 * Rmax, cb0 = 32 * Cmin;
    double accR[4][2], accC[4][2];
#pragma unroll
    for (int x = 0; x < 4; ++x) accR[x][0] = accR[x][1] = accC[x][0] = accC[x][1] = 0.0;
    if (diag) {
        // full symmetric tile from the stored lower triangle: element (r, c) = A[max][min]
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const int c8 = cb0 + 8 * cb;
            const double bv0 = S.V[g * ld + c8 + t], bv1 = S.V[g * ld + c8 + 4 + t];
            double a0[4], a1[4];
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) {
                const int r = rb0 + 8 * rb + g, c = c8 + t, c2 = c8 + 4 + t;
                const int hi = max(r, c), lo = min(r, c), hi2 = max(r, c2), lo2 = min(r, c2);
                a0[rb] = (hi < m) ? A22[(size_t)lo * lda + hi] : 0.0;
                a1[rb] = (hi2 < m) ? A22[(size_t)lo2 * lda + hi2] : 0.0;
            }
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) {
                dmma884(accR[rb][0], accR[rb][1], a0[rb], bv0);
                dmma884(accR[rb][0], accR[rb][1], a1[rb], bv1);
            }
        }
    } else {
        // software pipeline over the 4 column blocks: the 16 global fragments of block cb+1 are in
        // flight while the 16 DMMAs of block cb execute
        double a0[4], a1[4], s0[4], s1[4];
        auto load_cb = [&](int cb) {
            const int c8 = cb0 + 8 * cb;
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) {
                const int r8 = rb0 + 8 * rb, r = r8 + g, rr0 = r8 + t, rr1 = r8 + 4 + t;
                // contribution 1: A(g, k) = S[g][4ks + t];  contribution 2: A(g, k) = S^T[g][4ks + t] = S[4ks + t][g]
                a0[rb] = (r < m) ? A22[(size_t)(c8 + t) * lda + r] : 0.0;
                a1[rb] = (r < m) ? A22[(size_t)(c8 + 4 + t) * lda + r] : 0.0;
                s0[rb] = (rr0 < m) ? A22[(size_t)(c8 + g) * lda + rr0] : 0.0;
                s1[rb] = (rr1 < m) ? A22[(size_t)(c8 + g) * lda + rr1] : 0.0;
            }
        };
        load_cb(0);
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const int c8 = cb0 + 8 * cb;
            const double bv0 = S.V[g * ld + c8 + t], bv1 = S.V[g * ld + c8 + 4 + t];
            double ca0[4], ca1[4], cs0[4], cs1[4];
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) { ca0[rb] = a0[rb]; ca1[rb] = a1[rb]; cs0[rb] = s0[rb]; cs1[rb] = s1[rb]; }
            if (cb < 3) load_cb(cb + 1);
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) {
                const int r8 = rb0 + 8 * rb;
                dmma884(accR[rb][0], accR[rb][1], ca0[rb], bv0);
                dmma884(accR[rb][0], accR[rb][1], ca1[rb], bv1);
                // B(k, n) = V[r8 + 4ks + t][n = g]
                dmma884(accC[cb][0], accC[cb][1], cs0[rb], S.V[g * ld + r8 + t]);
                dmma884(accC[cb][0], accC[cb][1], cs1[rb], S.V[g * ld + r8 + 4 + t]);
            }
        }
    }
    // accumulators: (row = 8x + g, col = 2t + h); own rows go to Y, partner rows to Y2
    double* ownR = (diag || a == Rmax) ? S.Y : S.Y2;
    double* ownC = (a == Rmax) ? S.Y2 : S.Y;
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            ownR[(2 * t + h) * ld + rb0 + 8 * x + g] += accR[x][h];
            if (!diag) ownC[(2 * t + h) * ld + cb0 + 8 * x + g] += accC[x][h];
        }
}

__global__ void __launch_bounds__(512, 1)
sy2sb_small_kernel(double* __restrict__ A_all, int N, double* __restrict__ AB_all) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x, lda = N;
    double* A = A_all + (size_t)b * N * N;
    double* AB = AB_all + (size_t)b * (NB + 1) * N;
    // leading dimension of the shared panels: >= rows rounded to 32, and == 8 (mod 16) for conflict-free fragment loads
    const int mp = ((N + 31) / 32) * 32;
    s1_smem S;
    S.ld = mp + 4;  // == 4 (mod 16): both DMMA fragment patterns are 2-way (optimal)
    S.V = smem;
    S.Y = S.V + 8 * S.ld;
    S.Y2 = S.Y + 8 * S.ld;
    S.G = S.Y2 + 8 * S.ld;
    S.Tm = S.G + 64;
    S.M2 = S.Tm + 64;
    S.Rs = S.M2 + 64;
    S.tau = S.Rs + 64;
    S.red = S.tau + 8;
    const int ld = S.ld;

    for (int k0 = 0; k0 < N; k0 += NB) {
        const int r0 = k0 + NB, m = N - r0;
        // ---- emit the (final) diagonal block k0 into band storage ----
        if (tid < 64) {
            const int j = tid & 7, dd = tid >> 3;
            if (j + dd < NB && k0 + j + dd < N) AB[(size_t)dd * N + k0 + j] = A[(size_t)(k0 + j) * lda + k0 + j + dd];
        }
        if (m <= 0) break;
        double* A22 = A + (size_t)r0 * lda + r0;
        // ---- 1. panel -> shared, Householder QR ----
        for (int idx = tid; idx < 8 * ld; idx += T) {
            const int j = idx / ld, i = idx % ld;
            S.V[idx] = (i < m) ? A[(size_t)(k0 + j) * lda + r0 + i] : 0.0;
            S.Y[idx] = 0.0;
            S.Y2[idx] = 0.0;
        }
        if (tid < 64) S.Rs[tid] = 0.0;
        __syncthreads();
        const int nref = min(NB, m);
        for (int j = 0; j < NB; ++j) {
            double tau = 0.0;
            if (j < nref && m - j >= 2) {
                double part = 0.0;
                for (int i = j + 1 + tid; i < m; i += T) part = fma(S.V[j * ld + i], S.V[j * ld + i], part);
                const double tail2 = block_sum(part, S.red);
                const double x0 = S.V[j * ld + j];
                double beta = x0, inv = 0.0;
                if (tail2 > DBL_MIN) {
                    beta = sqrt(fma(x0, x0, tail2));
                    if (x0 >= 0.0) beta = -beta;
                    inv = 1.0 / (x0 - beta);
                    tau = (beta - x0) / beta;
                }
                __syncthreads();  // everyone has read x0 before it is overwritten
                // scale the reflector and form w_c = tau * (P[j][c] + sum_{i>j} v_i P[i][c]) for c > j
                double pw[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) pw[c] = 0.0;
                for (int i = j + 1 + tid; i < m; i += T) {
                    const double vi = S.V[j * ld + i] * inv;
                    S.V[j * ld + i] = vi;
#pragma unroll
                    for (int c = 0; c < NB; ++c)
                        if (c > j) pw[c] = fma(vi, S.V[c * ld + i], pw[c]);
                }
                block_sum_vec<NB>(pw, S.red, S.G);
                for (int i = j + tid; i < m; i += T) {
                    const double vi = (i == j) ? 1.0 : S.V[j * ld + i];
#pragma unroll
                    for (int c = 0; c < NB; ++c)
                        if (c > j) {
                            const double w = tau * (S.V[c * ld + j] + S.G[c]);
                            if (i != j) S.V[c * ld + i] = fma(-vi, w, S.V[c * ld + i]);
                        }
                }
                __syncthreads();
                if (tid < NB && tid > j) S.V[tid * ld + j] -= tau * (S.V[tid * ld + j] + S.G[tid]);  // row j itself (v_j = 1)
                if (tid == 0) S.V[j * ld + j] = beta;
                __syncthreads();
            }
            if (tid == 0) S.tau[j] = tau;
        }
        __syncthreads();
        // R (upper triangle of the top 8x8) -> Rs and band storage; then make V explicit (unit lower trapezoidal)
        if (tid < 64) {
            const int c = tid >> 3, i = tid & 7;
            if (i <= c && i < m) {
                const double r = S.V[c * ld + i];
                AB[(size_t)(NB + i - c) * N + k0 + c] = r;
            }
        }
        __syncthreads();
        if (tid < 64) {
            const int c = tid >> 3, i = tid & 7;
            if (i < c) S.V[c * ld + i] = 0.0;
            else if (i == c) S.V[c * ld + i] = (i < m) ? 1.0 : 0.0;
        }
        __syncthreads();
        bool any = false;
        for (int j = 0; j < NB; ++j) any = any || (S.tau[j] != 0.0);
        if (!any) continue;  // nothing to apply (uniform across the block)
        // ---- 2. T from the Gram matrix ----
        gram8(S.V, S.V, ld, m, S.red, S.G);
        if (tid == 0) {
            for (int j = 0; j < NB; ++j) {
                const double tj = S.tau[j];
                for (int a = 0; a < NB; ++a) S.Tm[a * 8 + j] = 0.0;
                S.Tm[j * 8 + j] = tj;
                // T(0:j, j) = -tau_j * T(0:j,0:j) * G(0:j, j)
                for (int a = 0; a < j; ++a) {
                    double s = 0.0;
                    for (int c = a; c < j; ++c) s += S.Tm[a * 8 + c] * S.G[c * 8 + j];
                    S.Tm[a * 8 + j] = -tj * s;
                }
            }
        }
        __syncthreads();
        // ---- 3. Y0 = A22 V (lower triangle stored): cyclic pairing of 32x32 tile tasks ----
        const int nt = (m + 31) >> 5;
        for (int wt = warp; wt < nt; wt += nwarps) symm_tile(A22, lda, m, wt, wt, S, lane);
        for (int s = 1; s <= (nt >> 1); ++s) {
            __syncthreads();
            const int lim = (2 * s == nt) ? (nt >> 1) : nt;
            for (int wt = warp; wt < lim; wt += nwarps) {
                int pt = wt - s;
                if (pt < 0) pt += nt;
                symm_tile(A22, lda, m, wt, pt, S, lane);
            }
        }
        __syncthreads();
        // ---- 4. Y = Y0 T;  X = V^T Y;  Z = Y - 1/2 V (T^T X) ----
        for (int i = tid; i < m; i += T) {
            double y0[NB], y[NB];
#pragma unroll
            for (int a = 0; a < NB; ++a) y0[a] = S.Y[a * ld + i] + S.Y2[a * ld + i];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < NB; ++a)
                    if (a <= c) s = fma(y0[a], S.Tm[a * 8 + c], s);
                y[c] = s;
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) S.Y[c * ld + i] = y[c];
        }
        __syncthreads();
        gram8(S.V, S.Y, ld, m, S.red, S.G);  // G = X = V^T Y
        if (tid < 64) {
            const int a = tid >> 3, c = tid & 7;  // M2 = T^T X
            double s = 0.0;
            for (int q = 0; q <= a; ++q) s += S.Tm[q * 8 + a] * S.G[q * 8 + c];
            S.M2[a * 8 + c] = s;
        }
        __syncthreads();
        for (int i = tid; i < m; i += T) {
            double v[NB];
#pragma unroll
            for (int a = 0; a < NB; ++a) v[a] = S.V[a * ld + i];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < NB; ++a) s = fma(v[a], S.M2[a * 8 + c], s);
                S.Y[c * ld + i] -= 0.5 * s;
            }
        }
        __syncthreads();
        // ---- 5. A22 -= V Z^T + Z V^T on the lower triangle (Z lives in S.Y) ----
        const int ntl = nt * (nt + 1) / 2;
        for (int tt = warp; tt < ntl; tt += nwarps) {
            int R = (int)((sqrtf(8.0f * (float)tt + 1.0f) - 1.0f) * 0.5f);
            while (R * (R + 1) / 2 > tt) --R;
            while ((R + 1) * (R + 2) / 2 <= tt) ++R;
            const int C = tt - R * (R + 1) / 2;
            const int rb0 = 32 * R, cb0 = 32 * C;
            // the C tile is loaded straight into the accumulators (all 32 loads in flight at once) and the
            // update is accumulated with a negated A operand: acc = C - [V Z][Z V]^T
            double acc[4][4][2];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int rr = rb0 + 8 * x + g, cc = cb0 + 8 * y + 2 * t + h;
                        acc[x][y][h] = (rr < m && cc < m && rr >= cc) ? A22[(size_t)cc * lda + rr] : 0.0;
                    }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                // P = [V Z] (A operand), Q = [Z V] (B operand); k = 4 ks + t
                const double* Pp = (ks < 2) ? S.V : S.Y;
                const double* Qp = (ks < 2) ? S.Y : S.V;
                const int col = 4 * (ks & 1) + t;
                double af[4], bf[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    af[x] = -Pp[col * ld + rb0 + 8 * x + g];
                    bf[x] = Qp[col * ld + cb0 + 8 * x + g];
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
                        const int rr = rb0 + 8 * x + g, cc = cb0 + 8 * y + 2 * t + h;
                        if (rr < m && cc < m && rr >= cc) A22[(size_t)cc * lda + rr] = acc[x][y][h];
                    }
        }
        __syncthreads();
    }
}

}  // namespace

size_t fkmc_sy2sb_small_smem(int N) {
    const size_t ld = ((N + 31) / 32) * 32 + 4;
    return sizeof(double) * (3 * 8 * ld + 4 * 64 + 8 + 16 * 64 + 72);
}

int fkmc_launch_sy2sb_small(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_AB) {
    fkmc_prof_scope ps(ctx, "sy2sb");
    const size_t smem = fkmc_sy2sb_small_smem(N);
    if (smem > ctx->smem_optin) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sy2sb: matrix too large for shared memory");
    int nwarps = (N + 31) / 32;
    if (nwarps < 2) nwarps = 2;
    if (nwarps > 16) nwarps = 16;
    FKMC_CUDA(ctx, cudaFuncSetAttribute(sy2sb_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sy2sb_small_kernel<<<B, nwarps * 32, smem, ctx->stream>>>(d_A, N, d_AB);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}
