// Stage 2 of the two-stage tridiagonalisation: symmetric band (half-bandwidth 8) -> tridiagonal by
// Householder bulge chasing, entirely in shared memory, one CTA per matrix.
//
// Together with sy2sb.cu this replaces the tridiagonalisation stage of Eigen::SelfAdjointEigenSolver
// as called from configuration_t::calc_ed (src/configuration.cpp:212-213).
//
// Storage: Wb[c][d] = A(c+d, c), d = 0..15 (band 0..8 plus room for the transient fill 9..15), i.e. each
// matrix column from its diagonal downward in 16 consecutive doubles.  Sweep j annihilates column j below
// the sub-diagonal with an 8-row reflector and chases the resulting bulge down the band in blocks of 8
// (Murata-Horikoshi / Lang).  Each warp owns whole sweeps; sweep j+1 may execute its step s once sweep j
// has finished step s+2, which is tracked with per-sweep progress counters in shared memory, so up to
// nwarps sweeps are in flight along the band.
#include <cfloat>

#include "common.cuh"

namespace {

#ifndef FKMC_SB2ST_SLEEP
#define FKMC_SB2ST_SLEEP 20
#endif
constexpr int SB = 8;     // half bandwidth
constexpr int WD = 16;    // stored sub-diagonals per column

struct refl {
    double v;    // lane i (0..7 within every group of 8) holds v_i; v_0 = 1
    double tau;
    double beta;
};

// sum over the 8 lanes that share (lane >> 3)
__device__ __forceinline__ double sum8(double x) {
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    return x;
}
// sum over the 4 groups (lanes with equal lane & 7)
__device__ __forceinline__ double sum4g(double x) {
    x += __shfl_xor_sync(0xffffffffu, x, 8);
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    return x;
}

// Householder reflector for x (x_i in lane i of each group, i < n; zero beyond n)
__device__ __forceinline__ refl make_reflector(double x, int i, int n) {
    refl R;
    const double tail2 = sum8((i >= 1 && i < n) ? x * x : 0.0);
    const double x0 = __shfl_sync(0xffffffffu, x, (threadIdx.x & 24));  // lane 0 of this group
    if (tail2 <= DBL_MIN) {
        R.tau = 0.0;
        R.beta = x0;
        R.v = (i == 0) ? 1.0 : 0.0;
    } else {
        double beta = sqrt(fma(x0, x0, tail2));
        if (x0 >= 0.0) beta = -beta;
        R.beta = beta;
        R.tau = (beta - x0) / beta;
        const double inv = 1.0 / (x0 - beta);
        R.v = (i == 0) ? 1.0 : ((i < n) ? x * inv : 0.0);
    }
    return R;
}

__global__ void __launch_bounds__(1024, 1)
sb2st_kernel(const double* __restrict__ AB_all, int N, double* __restrict__ d_all, double* __restrict__ e_all) {
    extern __shared__ double smem[];
    double* Wb = smem;                                         // [N + 16][16]
    volatile int* prog = reinterpret_cast<volatile int*>(Wb + (size_t)(N + 16) * WD);  // [N] steps completed per sweep
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const double* AB = AB_all + (size_t)b * (SB + 1) * N;

    for (int idx = tid; idx < (N + 16) * WD; idx += T) {
        const int c = idx / WD, dd = idx % WD;
        Wb[idx] = (c < N && dd <= SB && c + dd < N) ? AB[(size_t)dd * N + c] : 0.0;
    }
    for (int j = tid; j < N; j += T) prog[j] = 0;
    __syncthreads();

    const int i = lane & 7, q = lane >> 3;  // i: row index inside a block of 8, q: column group
    const int nsweeps = N - 2;
    for (int j = warp; j < nsweeps; j += nwarps) {
        int p = j + 1;                  // first row of the current reflector's index set
        int n = min(SB, N - p);         // its size
        // ---- step 0: annihilate column j below the sub-diagonal ----
        if (j > 0) {
            while (prog[j - 1] < 3) { __nanosleep(FKMC_SB2ST_SLEEP); }
            __threadfence_block();
        }
        double x = Wb[(size_t)j * WD + 1 + i];  // rows p..p+7 of column j (zero beyond the matrix)
        refl R = make_reflector(x, i, n);
        if (q == 0 && i < n) Wb[(size_t)j * WD + 1 + i] = (i == 0) ? R.beta : 0.0;
        int s = 0;
        while (true) {
            const double vi = R.v, tau = R.tau;
            // (i) [s >= 1] left-apply H to the other 7 columns of the bulge block A(J, p-8+1 .. p-1)
            if (s >= 1 && tau != 0.0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int cc = 1 + q + 4 * h;  // column offset inside the previous block, 1..8 (8 is out of range)
                    const bool ok = cc < SB;
                    const int c = p - SB + (ok ? cc : 1);
                    double* col = Wb + (size_t)c * WD;
                    const double a = ok ? col[p - c + i] : 0.0;
                    const double w = tau * sum8(vi * a);
                    if (ok) col[p - c + i] = a - vi * w;
                }
            }
            // (ii) two-sided update of the diagonal block D = A(J, J): lane (i, q) owns D(i, c) for c = q, q+4
            if (tau != 0.0) {
                double dcol[2], vc[2];
                double part = 0.0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = q + 4 * h;
                    const int hi = max(i, c), lo = min(i, c);
                    dcol[h] = Wb[(size_t)(p + lo) * WD + (hi - lo)];
                    vc[h] = __shfl_sync(0xffffffffu, vi, (lane & 24) + c);
                    part = fma(dcol[h], vc[h], part);
                }
                double u = tau * sum4g(part);                  // u_i = tau * (D v)_i
                const double alpha = -0.5 * tau * sum8(u * vi);
                u = fma(alpha, vi, u);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = q + 4 * h;
                    const double uc = __shfl_sync(0xffffffffu, u, (lane & 24) + c);
                    if (i >= c) Wb[(size_t)(p + c) * WD + (i - c)] = dcol[h] - vi * uc - u * vc[h];
                }
            }
            // (iii) right-apply H to the block below: A(Jn, J), Jn = rows p+8 .. p+15; lane (i, q) owns rows i, columns q, q+4
            const int pn = p + SB;
            const int nn = min(SB, N - pn);
            double xnext = 0.0;
            if (n == SB && nn > 0) {
                double bel[2], vc[2];
                double part = 0.0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = q + 4 * h;
                    bel[h] = Wb[(size_t)(p + c) * WD + (SB + i - c)];
                    vc[h] = __shfl_sync(0xffffffffu, vi, (lane & 24) + c);
                    part = fma(bel[h], vc[h], part);
                }
                const double z = tau * sum4g(part);            // z_i = tau * sum_c B(i, c) v_c
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = q + 4 * h;
                    bel[h] = fma(-z, vc[h], bel[h]);
                    if (tau != 0.0) Wb[(size_t)(p + c) * WD + (SB + i - c)] = bel[h];
                }
                xnext = __shfl_sync(0xffffffffu, bel[0], i);   // column 0 of the block lives in group q = 0, h = 0
            }
            __syncwarp();
            // publish progress, then move to the next block
            __threadfence_block();
            ++s;
            if (lane == 0) prog[j] = s;
            if (!(n == SB && nn >= 2)) break;
            if (j > 0) {
                while (prog[j - 1] < s + 3) { __nanosleep(FKMC_SB2ST_SLEEP); }
                __threadfence_block();
            }
            // reflector from the first column of the block below (rows pn.., column p); re-read after the wait is not
            // needed: that column is only ever touched by this sweep at this point
            R = make_reflector(xnext, i, nn);
            if (q == 0 && i < nn) Wb[(size_t)p * WD + SB + i] = (i == 0) ? R.beta : 0.0;
            p = pn;
            n = nn;
        }
        __threadfence_block();
        if (lane == 0) prog[j] = 1 << 30;
    }
    __syncthreads();
    double* dd = d_all + (size_t)b * N;
    double* ee = e_all + (size_t)b * N;
    for (int c = tid; c < N; c += T) {
        dd[c] = Wb[(size_t)c * WD];
        ee[c] = (c < N - 1) ? Wb[(size_t)c * WD + 1] : 0.0;
    }
}

}  // namespace

size_t fkmc_sb2st_smem(int N) { return sizeof(double) * (size_t)(N + 16) * WD + sizeof(int) * (size_t)N + 16; }

int fkmc_launch_sb2st(fkmc_ctx* ctx, const double* d_AB, int N, int B, double* d_d, double* d_e) {
    fkmc_prof_scope ps(ctx, "sb2st");
    const size_t smem = fkmc_sb2st_smem(N);
    if (smem > ctx->smem_optin) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sb2st: matrix too large for shared memory");
    int nwarps = (N + 23) / 24;  // ~one warp per 3 blocks of the band (the pipeline lag)
    if (nwarps < 1) nwarps = 1;
    if (nwarps > ((N > 640) ? 32 : 16)) nwarps = (N > 640) ? 32 : 16;  // <= 512 threads keeps two CTAs per SM for N <= 512
    FKMC_CUDA(ctx, cudaFuncSetAttribute(sb2st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sb2st_kernel<<<B, nwarps * 32, smem, ctx->stream>>>(d_AB, N, d_d, d_e);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}
