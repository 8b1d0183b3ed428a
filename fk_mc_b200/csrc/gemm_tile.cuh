// Batched FP64 GEMM tile on the tensor cores (mma.sync m8n8k4 -> DMMA.8x8x4), shared by the eigenvector update of the fast-update path
// (secular.cu) and the current-current contraction of the stiffness measure (stiffness.cu).
//   C[i][j] = sum_k A[i][k] B[k][j],  all row-major N x N.
// CTA tile 128 x 64, 256 threads = 8 warps of 32 x 32 (16 DMMA accumulators each), K in chunks of 32 double-buffered in shared memory
// through 16-byte asynchronous copies (LDGSTS); two CTAs per SM (108.5 KB of shared memory each).
#pragma once
#include "common.cuh"

namespace fkgemm {

// DMMA without the volatile qualifier (a pure function of its operands): ptxas may reorder it against the fragment loads
__device__ __forceinline__ void dmma_nv(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int GM = 128, GN = 64, GK = 32;
constexpr int SA = GK + 4;   // row stride of the A chunk (== 4 mod 16: conflict-free fragment reads)
constexpr int SB = GN + 4;   // row stride of the B chunk


constexpr size_t smem_bytes() { return sizeof(double) * (2 * GM * SA + 2 * GK * SB); }
__host__ __device__ inline int tiles(int N) { return ((N + GM - 1) / GM) * ((N + GN - 1) / GN); }

// one CTA tile: tile index -> (i0, j0); sm = dynamic shared memory (smem_bytes(), 16-byte aligned)
__device__ __forceinline__ void tile(const double* __restrict__ A, const double* __restrict__ Bq, double* __restrict__ Cm, int N, int tile_index, double* sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int ntn = (N + GN - 1) / GN;
    const int i0 = (tile_index / ntn) * GM, j0 = (tile_index % ntn) * GN;
    double* As = sm;                 // [2][GM][SA]
    double* Bs = As + 2 * GM * SA;   // [2][GK][SB]
    const int wm = warp >> 1, wn = warp & 1;  // warp tile: rows 32 wm .. +31, cols 32 wn .. +31
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int nchunk = (N + GK - 1) / GK;
    // A chunk: 128 rows x 32 doubles; thread loads row (tid >> 1), 16 doubles starting at 16 (tid & 1)
    // B chunk: 32 rows x 64 doubles; thread loads row (tid >> 3), 8 doubles starting at 8 (tid & 7)
    const int ar = tid >> 1, ac = (tid & 1) * 16;
    const int br = tid >> 3, bc = (tid & 7) * 8;
    const bool vec_ok = (N % 2 == 0);
    // 16-byte asynchronous copies global -> shared (LDGSTS), zero-filled past the matrix edge; odd N takes a synchronous scalar path
    auto copy2 = [&](double* dstp, const double* M, int row, int col) {
        if (vec_ok) {
            const int nb = (row < N && col < N) ? 16 : 0;   // N even and col even: a pair is either inside or outside
            const double* srcp = nb ? M + (size_t)row * N + col : M;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dstp)), "l"(srcp), "r"(nb) : "memory");
        } else {
            dstp[0] = (row < N && col < N) ? M[(size_t)row * N + col] : 0.0;
            dstp[1] = (row < N && col + 1 < N) ? M[(size_t)row * N + col + 1] : 0.0;
        }
    };
    auto issue = [&](int ch, int buf) {
        double* ap = As + (size_t)buf * GM * SA + ar * SA + ac;
#pragma unroll
        for (int u = 0; u < 16; u += 2) copy2(ap + u, A, i0 + ar, ch * GK + ac + u);
        double* bp = Bs + (size_t)buf * GK * SB + br * SB + bc;
#pragma unroll
        for (int u = 0; u < 8; u += 2) copy2(bp + u, Bq, ch * GK + br, j0 + bc + u);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0, 0);
    for (int ch = 0; ch < nchunk; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunk) {
            issue(ch + 1, buf ^ 1);   // the other buffer was released by the barrier that ended the previous iteration
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* ap = As + (size_t)buf * GM * SA + (32 * wm) * SA;
        const double* bp = Bs + (size_t)buf * GK * SB + 32 * wn;
#pragma unroll
        for (int k4 = 0; k4 < GK; k4 += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = ap[(8 * a + g) * SA + k4 + t];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = bp[(k4 + t) * SB + 8 * b + g];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma_nv(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gi = i0 + 32 * wm + 8 * a + g, gj = j0 + 32 * wn + 8 * b + 2 * t;
            if (gi < N) {
                if (vec_ok && gj + 1 < N) *reinterpret_cast<double2*>(Cm + (size_t)gi * N + gj) = make_double2(acc[a][b][0], acc[a][b][1]);
                else {
                    if (gj < N) Cm[(size_t)gi * N + gj] = acc[a][b][0];
                    if (gj + 1 < N) Cm[(size_t)gi * N + gj + 1] = acc[a][b][1];
                }
            }
        }
}

}  // namespace fkgemm
