// Stage 1 of the two-stage tridiagonalisation: dense symmetric (lower) -> symmetric band with
// half-bandwidth 8, one CTA (8 warps) per matrix, every O(N^3) flop on the FP64 tensor cores.
//
// Together with sb2st.cu this replaces the tridiagonalisation stage of Eigen::SelfAdjointEigenSolver
// as called from configuration_t::calc_ed (src/configuration.cpp:212-213).  For block column k
// (columns k0..k0+7, trailing rows r0 = k0+8 .. N-1, m = N - r0):
//   1. Householder QR of the m x 8 panel in shared memory  ->  V (unit lower trapezoidal), tau, R
//   2. T (8x8, compact WY: Q = I - V T V^T) from the Gram matrix V^T V (DMMA)
//   3. Y0 = A22 V          -- SYMM over the stored lower triangle in 32x32 tiles
//   4. Y = Y0 T,  X = V^T Y (DMMA),  Z = Y - 1/2 V (T^T X)
//   5. A22 -= V Z^T + Z V^T  -- rank-16 SYR2K on the lower triangle
// Steps 3 and 5 stream the trailing matrix tile by tile: every warp owns a 32x32 shared-memory tile
// buffer that is filled by bulk asynchronous copies (cp.async.bulk + mbarrier transaction counts, one
// 256-byte column per lane -- SASS UBLKCP) and, in step 5, drained by bulk stores, so the tensor-core
// warps never hold global loads in registers.  DMMA fragments are read from the tile with a column
// stride == 4 (mod 16) doubles, which is conflict-free for both the direct and the transposed operand.
// In step 3 the tile tasks are paired cyclically ({a, a-s}, s = 0..nt/2): a warp keeps the rows of
// its own tiles in registers for the whole pass and adds the partner rows straight into shared
// memory -- in one step all partner tiles are distinct, so no atomics are needed.
// R and the diagonal blocks are emitted in band storage AB[d][c] = A(c+d, c), d = 0..8.
// Matrix layout ("tiled"): only the lower-triangular 32x32 tiles are stored, tile (R, C), R >= C, at offset
// (R(R+1)/2 + C) * 1152 doubles, column-major inside the tile with the same padded column stride of 36 doubles that the
// shared-memory copy uses -- so a tile moves with ONE 9216-byte bulk copy in each direction and lands bank-conflict-free.
// The tile grid is fixed in global coordinates; the panels V, Y, Z are indexed by global row and are zero above the
// trailing block, which makes the partial edge tiles of each block column come out right without masks.
// Requires N % 8 == 0 and N <= 1024.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int NB = 8;
constexpr int NW = 8;       // warps per CTA
constexpr int TS = 36;      // tile column stride in doubles (== 4 mod 16)
constexpr int TILE = 32 * TS;
constexpr int MAXOWN = 4;   // owned row tiles per warp: N <= 1024 -> nt <= 32 -> 4

struct s1_smem {
    double* V;    // [8][ld] column-major panel / reflectors
    double* Y;    // [8][ld] Y0 -> Y -> Z
    double* G;    // [64] Gram / X / scratch
    double* Tm;   // [64] T
    double* M2;   // [64] T^T X
    double* tau;  // [8]
    double* red;  // [NW*64 + 72]
    double* tile; // [NW][TILE]
    uint64_t* bar; // [NW]
    int ld;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FKMC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FKMC_DONE_%=;\n"
        "bra FKMC_WAIT_%=;\n"
        "FKMC_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ size_t tile_off(int R, int C) { return ((size_t)R * (R + 1) / 2 + C) * TILE; }

// Issue the bulk load of global tile (R, C) into this warp's buffer (whole warp calls).
__device__ __forceinline__ void tile_load(double* buf, uint64_t* bar, const double* __restrict__ At, int R, int C, int lane) {
    __syncwarp();  // every lane is done reading the previous contents
    if (lane == 0) {
        mbar_expect_tx(bar, (uint32_t)(TILE * 8));
        bulk_g2s(buf, At + tile_off(R, C), (uint32_t)(TILE * 8), bar);
    }
}

// element (i, j), i >= j, of the tiled matrix
__device__ __forceinline__ size_t elem_off(int i, int j) { return tile_off(i >> 5, j >> 5) + (size_t)(j & 31) * TS + (i & 31); }

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

// SYMM on one staged tile.  Stored tile (Rmax, Cmin), S = tile contents:
//   accR (rows of Rmax) += S V[Cmin rows];   accC (rows of Cmin) += S^T V[Rmax rows]   (off-diagonal tiles)
//   accR += sym(S) V[rows]                                                            (diagonal tiles)
__device__ __forceinline__ void symm_tile(const double* __restrict__ Tb, bool diag, int rb0, int cb0, const double* __restrict__ V, int ld,
                                          int lane, double (&accR)[4][2], double (&accC)[4][2]) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
        const int c8 = 8 * cb;
        const double bv0 = V[g * ld + cb0 + c8 + t], bv1 = V[g * ld + cb0 + c8 + 4 + t];
#pragma unroll
        for (int rb = 0; rb < 4; ++rb) {
            const int r8 = 8 * rb;
            if (diag) {
                const int r = r8 + g, c = c8 + t, c2 = c8 + 4 + t;
                const double a0 = Tb[min(r, c) * TS + max(r, c)];
                const double a1 = Tb[min(r, c2) * TS + max(r, c2)];
                dmma884(accR[rb][0], accR[rb][1], a0, bv0);
                dmma884(accR[rb][0], accR[rb][1], a1, bv1);
            } else {
                // contribution 1: A(g, k) = S[g][4ks + t];  contribution 2: A(g, k) = S[4ks + t][g], B(k, n) = V[rb0 + r8 + 4ks + t][n = g]
                const double a0 = Tb[(c8 + t) * TS + r8 + g], a1 = Tb[(c8 + 4 + t) * TS + r8 + g];
                const double s0 = Tb[(c8 + g) * TS + r8 + t], s1 = Tb[(c8 + g) * TS + r8 + 4 + t];
                dmma884(accR[rb][0], accR[rb][1], a0, bv0);
                dmma884(accR[rb][0], accR[rb][1], a1, bv1);
                dmma884(accC[cb][0], accC[cb][1], s0, V[g * ld + rb0 + r8 + t]);
                dmma884(accC[cb][0], accC[cb][1], s1, V[g * ld + rb0 + r8 + 4 + t]);
            }
        }
    }
}

__global__ void __launch_bounds__(NW * 32, 1)
sy2sb_kernel(double* __restrict__ A_all, size_t a_stride, int N, double* __restrict__ AB_all) {
    extern __shared__ __align__(128) double smem[];
    const int tid = threadIdx.x, T = NW * 32, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x;
    double* At = A_all + (size_t)b * a_stride;  // tiled lower-triangular storage of this matrix
    double* AB = AB_all + (size_t)b * (NB + 1) * N;
    const int mp = ((N + 31) / 32) * 32, NT = mp >> 5;
    s1_smem S0;
    S0.ld = mp + 4;  // == 4 (mod 16): both fragment patterns are 2-way (optimal) on the panels
    S0.V = smem;
    S0.Y = S0.V + 8 * S0.ld;
    S0.G = S0.Y + 8 * S0.ld;
    S0.Tm = S0.G + 64;
    S0.M2 = S0.Tm + 64;
    S0.tau = S0.M2 + 64;
    S0.red = S0.tau + 8;
    S0.tile = S0.red + NW * 64 + 72;
    S0.bar = reinterpret_cast<uint64_t*>(S0.tile + NW * TILE);
    const int ld = S0.ld;
    double* Tb = S0.tile + warp * TILE;
    uint64_t* bar = S0.bar + warp;
    uint32_t parity = 0;

    for (int i = tid; i < NW * TILE; i += T) S0.tile[i] = 0.0;
    if (tid < NW) mbar_init(S0.bar + tid, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    for (int k0 = 0; k0 < N; k0 += NB) {
        const int r0 = k0 + NB, m = N - r0;
        const s1_smem SG = S0;  // panels indexed by global row (tile loops)
        s1_smem S = S0;         // panels indexed by local row i = global row - r0 (QR and the thin updates)
        S.V = S0.V + r0;
        S.Y = S0.Y + r0;
        // ---- emit the (final) diagonal block k0 into band storage ----
        if (tid < 64) {
            const int j = tid & 7, dd = tid >> 3;
            if (j + dd < NB && k0 + j + dd < N) AB[(size_t)dd * N + k0 + j] = At[elem_off(k0 + j + dd, k0 + j)];
        }
        if (m <= 0) break;
        // ---- 1. panel -> shared (global row index; zero above the trailing block), Householder QR ----
        for (int idx = tid; idx < 8 * ld; idx += T) {
            const int j = idx / ld, gi = idx % ld;
            SG.V[idx] = (gi >= r0 && gi < N) ? At[elem_off(gi, k0 + j)] : 0.0;
            SG.Y[idx] = 0.0;
        }
        __syncthreads();
        const int nref = min(NB, m);
        bool any = false;
        for (int j = 0; j < NB; ++j) {
            double tau = 0.0;
            if (j < nref && m - j >= 2) {
                // one fused reduction: pw[c] = sum_{i>j} P[i][j] P[i][c], c >= j  (c = j gives the tail norm)
                double pw[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) pw[c] = 0.0;
                for (int i = j + 1 + tid; i < m; i += T) {
                    const double pj = S.V[j * ld + i];
#pragma unroll
                    for (int c = 0; c < NB; ++c)
                        if (c >= j) pw[c] = fma(pj, S.V[c * ld + i], pw[c]);
                }
                block_sum_vec<NB>(pw, S.red, S.G);
                const double tail2 = S.G[j];
                const double x0 = S.V[j * ld + j];
                double beta = x0, inv = 0.0;
                if (tail2 > DBL_MIN) {
                    beta = sqrt(fma(x0, x0, tail2));
                    if (x0 >= 0.0) beta = -beta;
                    inv = 1.0 / (x0 - beta);
                    tau = (beta - x0) / beta;
                }
                // w_c = tau (P[j][c] + inv * G[c]);  rows i > j: v_i = P[i][j] inv, P[i][c] -= v_i w_c
                double w[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) w[c] = (c > j) ? tau * fma(inv, S.G[c], S.V[c * ld + j]) : 0.0;
                __syncthreads();  // all threads hold x0, G and row j before they change
                for (int i = j + 1 + tid; i < m; i += T) {
                    const double vi = S.V[j * ld + i] * inv;
                    S.V[j * ld + i] = vi;
#pragma unroll
                    for (int c = 0; c < NB; ++c)
                        if (c > j) S.V[c * ld + i] = fma(-vi, w[c], S.V[c * ld + i]);
                }
                if (tid < NB && tid > j) S.V[tid * ld + j] -= w[tid];  // row j itself (v_j = 1)
                if (tid == 0) S.V[j * ld + j] = beta;
                __syncthreads();
            }
            if (tid == 0) S.tau[j] = tau;
            any = any || (tau != 0.0);
        }
        __syncthreads();
        // R (upper triangle of the top 8x8) -> band storage; then make V explicit (unit lower trapezoidal)
        if (tid < 64) {
            const int c = tid >> 3, i = tid & 7;
            if (i <= c && i < m) AB[(size_t)(NB + i - c) * N + k0 + c] = S.V[c * ld + i];
        }
        __syncthreads();
        if (tid < 64) {
            const int c = tid >> 3, i = tid & 7;
            if (i < c) S.V[c * ld + i] = 0.0;
            else if (i == c) S.V[c * ld + i] = (i < m) ? 1.0 : 0.0;
        }
        __syncthreads();
        if (!any) continue;  // nothing to apply (uniform across the block)
        // ---- 2. T from the Gram matrix: T(0:j, j) = -tau_j T(0:j,0:j) G(0:j, j) ----
        gram8(S.V, S.V, ld, m, S.red, S.G);
        if (tid < 64) S.Tm[tid] = 0.0;
        __syncthreads();
        for (int j = 0; j < NB; ++j) {
            if (tid < j) {
                double s = 0.0;
                for (int c = tid; c < j; ++c) s += S.Tm[tid * 8 + c] * S.G[c * 8 + j];
                S.Tm[tid * 8 + j] = -S.tau[j] * s;
            } else if (tid == j) {
                S.Tm[j * 8 + j] = S.tau[j];
            }
            __syncthreads();
        }
        // ---- 3. Y0 = A22 V: cyclic pairing of tile tasks, own rows in registers, partner rows in shared memory ----
        const int T0 = r0 >> 5, nt = NT - T0;  // tiles T0..NT-1 of the fixed global grid touch the trailing block
        {
            double own[MAXOWN][4][2];
#pragma unroll
            for (int o = 0; o < MAXOWN; ++o)
#pragma unroll
                for (int x = 0; x < 4; ++x) own[o][x][0] = own[o][x][1] = 0.0;
            const int smax = nt >> 1;
            // prefetch the first task of this warp (its diagonal tile)
            if (warp < nt) tile_load(Tb, bar, At, T0 + warp, T0 + warp, lane);
            for (int s = 0; s <= smax; ++s) {
                const int lim = (s > 0 && 2 * s == nt) ? (nt >> 1) : nt;
#pragma unroll
                for (int o = 0; o < MAXOWN; ++o) {
                    const int a = warp + NW * o;
                    if (a < lim) {
                        int p = a - s;
                        if (p < 0) p += nt;
                        const bool diag = (s == 0);
                        const int Rmax = max(a, p), Cmin = min(a, p);
                        mbar_wait(bar, parity);
                        parity ^= 1;
                        double accR[4][2], accC[4][2];
#pragma unroll
                        for (int x = 0; x < 4; ++x) accR[x][0] = accR[x][1] = accC[x][0] = accC[x][1] = 0.0;
                        symm_tile(Tb, diag, 32 * (T0 + Rmax), 32 * (T0 + Cmin), SG.V, ld, lane, accR, accC);
                        // prefetch this warp's next task while the results are folded
                        {
                            int na = a + NW, ns = s;
                            const int nlim = lim;
                            if (na >= nlim) {
                                ns = s + 1;
                                na = warp;
                            }
                            const int nlim2 = (ns > 0 && 2 * ns == nt) ? (nt >> 1) : nt;
                            if (ns <= smax && na < nlim2) {
                                int np = na - ns;
                                if (np < 0) np += nt;
                                tile_load(Tb, bar, At, T0 + max(na, np), T0 + min(na, np), lane);
                            }
                        }
                        // own rows stay in registers; partner rows go to shared memory (distinct tiles within a step)
                        if (diag || a == Rmax) {
#pragma unroll
                            for (int x = 0; x < 4; ++x) { own[o][x][0] += accR[x][0]; own[o][x][1] += accR[x][1]; }
                            if (!diag) {
#pragma unroll
                                for (int x = 0; x < 4; ++x)
#pragma unroll
                                    for (int h = 0; h < 2; ++h) SG.Y[(2 * t + h) * ld + 32 * (T0 + Cmin) + 8 * x + g] += accC[x][h];
                            }
                        } else {
#pragma unroll
                            for (int x = 0; x < 4; ++x) { own[o][x][0] += accC[x][0]; own[o][x][1] += accC[x][1]; }
#pragma unroll
                            for (int x = 0; x < 4; ++x)
#pragma unroll
                                for (int h = 0; h < 2; ++h) SG.Y[(2 * t + h) * ld + 32 * (T0 + Rmax) + 8 * x + g] += accR[x][h];
                        }
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int o = 0; o < MAXOWN; ++o) {
                const int a = warp + NW * o;
                if (a < nt) {
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int h = 0; h < 2; ++h) SG.Y[(2 * t + h) * ld + 32 * (T0 + a) + 8 * x + g] += own[o][x][h];
                }
            }
        }
        __syncthreads();
        // rows of the first (edge) tile that lie above the trailing block picked up band entries: Z must be zero there
        for (int idx = tid; idx < 8 * (r0 - 32 * T0); idx += T) {
            const int j = idx / (r0 - 32 * T0), gi = 32 * T0 + idx % (r0 - 32 * T0);
            SG.Y[j * ld + gi] = 0.0;
        }
        // ---- 4. Y = Y0 T;  X = V^T Y;  Z = Y - 1/2 V (T^T X) ----
        for (int i = tid; i < m; i += T) {
            double y0[NB], y[NB];
#pragma unroll
            for (int a = 0; a < NB; ++a) y0[a] = S.Y[a * ld + i];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int a = 0; a < NB; ++a)
                    if (a <= c) sacc = fma(y0[a], S.Tm[a * 8 + c], sacc);
                y[c] = sacc;
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) S.Y[c * ld + i] = y[c];
        }
        __syncthreads();
        gram8(S.V, S.Y, ld, m, S.red, S.G);  // G = X = V^T Y
        if (tid < 64) {
            const int a = tid >> 3, c = tid & 7;  // M2 = T^T X
            double sacc = 0.0;
            for (int q = 0; q <= a; ++q) sacc += S.Tm[q * 8 + a] * S.G[q * 8 + c];
            S.M2[a * 8 + c] = sacc;
        }
        __syncthreads();
        for (int i = tid; i < m; i += T) {
            double v[NB];
#pragma unroll
            for (int a = 0; a < NB; ++a) v[a] = S.V[a * ld + i];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int a = 0; a < NB; ++a) sacc = fma(v[a], S.M2[a * 8 + c], sacc);
                S.Y[c * ld + i] -= 0.5 * sacc;
            }
        }
        __syncthreads();
        // ---- 5. A22 -= V Z^T + Z V^T on the lower triangle (Z lives in S.Y): load tile, update in shared memory, bulk store ----
        const int ntl = nt * (nt + 1) / 2;
        auto decode = [](int tt, int& R, int& C) {
            R = (int)((sqrtf(8.0f * (float)tt + 1.0f) - 1.0f) * 0.5f);
            while (R * (R + 1) / 2 > tt) --R;
            while ((R + 1) * (R + 2) / 2 <= tt) ++R;
            C = tt - R * (R + 1) / 2;
        };
        if (warp < ntl) {
            int R, C;
            decode(warp, R, C);
            tile_load(Tb, bar, At, T0 + R, T0 + C, lane);
        }
        for (int tt = warp; tt < ntl; tt += NW) {
            int R, C;
            decode(tt, R, C);
            const int rb0 = 32 * (T0 + R), cb0 = 32 * (T0 + C);  // global rows / columns
            // operand fragments do not depend on the tile contents: fetch them while the copy is in flight
            double af[4][4], bf[4][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                // P = [V Z] (A operand, negated), Q = [Z V] (B operand); k = 4 ks + t
                const double* Pp = (ks < 2) ? SG.V : SG.Y;
                const double* Qp = (ks < 2) ? SG.Y : SG.V;
                const int col = 4 * (ks & 1) + t;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    af[ks][x] = -Pp[col * ld + rb0 + 8 * x + g];
                    bf[ks][x] = Qp[col * ld + cb0 + 8 * x + g];
                }
            }
            mbar_wait(bar, parity);
            parity ^= 1;
            double acc[4][4][2];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int h = 0; h < 2; ++h) acc[x][y][h] = Tb[(8 * y + 2 * t + h) * TS + 8 * x + g];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) dmma884(acc[x][y][0], acc[x][y][1], af[ks][x], bf[ks][y]);
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
#pragma unroll
                    for (int h = 0; h < 2; ++h) Tb[(8 * y + 2 * t + h) * TS + 8 * x + g] = acc[x][y][h];
            // drain: generic-proxy writes -> async proxy, then one bulk store of the whole (padded) tile.  Entries outside the
            // trailing block see a zero update (V, Z are zero there); the strictly upper part of a diagonal tile is never read.
            fence_async_smem();
            __syncwarp();
            if (lane == 0) bulk_s2g(At + tile_off(T0 + R, T0 + C), Tb, (uint32_t)(TILE * 8));
            bulk_commit();
            bulk_wait_read();  // the buffer may be overwritten once the store has read it
            const int tn = tt + NW;
            if (tn < ntl) {
                int R2, C2;
                decode(tn, R2, C2);
                tile_load(Tb, bar, At, T0 + R2, T0 + C2, lane);
            }
        }
        // all bulk stores of this block column must have landed before the next panel is read
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncthreads();
    }
}

}  // namespace

size_t fkmc_sy2sb_smem(int N) {
    const size_t ld = ((N + 31) / 32) * 32 + 4;
    return sizeof(double) * (2 * 8 * ld + 3 * 64 + 8 + NW * 64 + 72 + (size_t)NW * TILE) + sizeof(uint64_t) * NW + 16;
}

int fkmc_launch_sy2sb_small(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_AB);

// doubles per matrix in the tiled lower-triangular layout
size_t fkmc_tiled_stride(int N) {
    const size_t nt = (N + 31) / 32;
    return nt * (nt + 1) / 2 * TILE;
}
bool fkmc_use_tiled(int N) { return N >= 512 && N % 8 == 0 && N <= 1024; }

// dense -> band on a batch of matrices in the tiled layout (see fkmc_launch_build_h_tiled / fkmc_launch_to_tiled)
int fkmc_launch_sy2sb_tiled(fkmc_ctx* ctx, double* d_At, int N, int B, double* d_AB) {
    fkmc_prof_scope ps(ctx, "sy2sb");
    if (!fkmc_use_tiled(N)) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sy2sb_tiled: needs 512 <= N <= 1024, N % 8 == 0");
    const size_t smem = fkmc_sy2sb_smem(N);
    if (smem > ctx->smem_optin) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sy2sb: matrix too large for shared memory");
    FKMC_CUDA(ctx, cudaFuncSetAttribute(sy2sb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sy2sb_kernel<<<B, NW * 32, smem, ctx->stream>>>(d_At, fkmc_tiled_stride(N), N, d_AB);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

// column-major [B][N][N] (lower triangle) -> tiled layout
__global__ void __launch_bounds__(256) to_tiled_kernel(const double* __restrict__ A_all, int N, double* __restrict__ At_all, size_t t_stride) {
    const int b = blockIdx.y, tile = blockIdx.x;
    int R = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
    while (R * (R + 1) / 2 > tile) --R;
    while ((R + 1) * (R + 2) / 2 <= tile) ++R;
    const int C = tile - R * (R + 1) / 2;
    const double* A = A_all + (size_t)b * N * N;
    double* dst = At_all + (size_t)b * t_stride + (size_t)tile * TILE;
    for (int e = threadIdx.x; e < TILE; e += blockDim.x) {
        const int c = e / TS, r = e % TS;
        const int i = 32 * R + r, j = 32 * C + c;
        double v = 0.0;
        if (r < 32 && i < N && j < N) v = A[(size_t)min(i, j) * N + max(i, j)];  // symmetric fill from the lower triangle
        dst[e] = v;
    }
}

int fkmc_launch_to_tiled(fkmc_ctx* ctx, const double* d_A, int N, int B, double* d_At) {
    const int nt = (N + 31) / 32;
    dim3 grid(nt * (nt + 1) / 2, B);
    to_tiled_kernel<<<grid, 256, 0, ctx->stream>>>(d_A, N, d_At, fkmc_tiled_stride(N));
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

// Hamiltonian assembly straight into the tiled layout: H = hopping + diag(U f - mu_c)
// (configuration_t::calc_hamiltonian, src/configuration.cpp:79-91)
__global__ void __launch_bounds__(256) build_h_tiled_kernel(const int32_t* __restrict__ f, const int* __restrict__ nbr_idx,
                                                            const double* __restrict__ nbr_val, int N, int Z, double U, double mu_c,
                                                            double* __restrict__ At_all, size_t t_stride) {
    const int b = blockIdx.y, tile = blockIdx.x;
    int R = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
    while (R * (R + 1) / 2 > tile) --R;
    while ((R + 1) * (R + 2) / 2 <= tile) ++R;
    const int C = tile - R * (R + 1) / 2;
    const int32_t* fb = f + (size_t)b * N;
    double* dst = At_all + (size_t)b * t_stride + (size_t)tile * TILE;
    for (int e = threadIdx.x; e < TILE; e += blockDim.x) {
        const int c = e / TS, r = e % TS;
        const int i = 32 * R + r, j = 32 * C + c;
        double v = 0.0;
        if (r < 32 && i < N && j < N) {
            if (i == j) {
                v = U * (double)fb[i] - mu_c;
            } else {
                for (int z = 0; z < Z; ++z)
                    if (nbr_idx[z * N + i] == j) v = nbr_val[z * N + i];
            }
        }
        dst[e] = v;
    }
}

int fkmc_launch_build_h_tiled(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double* d_At) {
    fkmc_prof_scope ps(ctx, "build_h");
    const int nt = (ctx->N + 31) / 32;
    dim3 grid(nt * (nt + 1) / 2, B);
    build_h_tiled_kernel<<<grid, 256, 0, ctx->stream>>>(d_f, ctx->d_nbr_idx, ctx->d_nbr_val, ctx->N, ctx->Z, U, mu_c, d_At,
                                                       fkmc_tiled_stride(ctx->N));
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

// column-major entry point (stage-level API and small matrices)
int fkmc_launch_sy2sb(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_AB) {
    if (!fkmc_use_tiled(N)) return fkmc_launch_sy2sb_small(ctx, d_A, N, B, d_AB);
    // convert in place is not possible (layouts overlap): stage through a scratch allocation
    double* d_At = nullptr;
    FKMC_CUDA(ctx, cudaMalloc(&d_At, sizeof(double) * fkmc_tiled_stride(N) * (size_t)B));
    int rc = fkmc_launch_to_tiled(ctx, d_A, N, B, d_At);
    if (!rc) rc = fkmc_launch_sy2sb_tiled(ctx, d_At, N, B, d_AB);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_At);
    return rc;
}
