// Stage 1 of the two-stage tridiagonalisation: dense symmetric (lower) -> symmetric band with
// half-bandwidth 8, one CTA (8 warps) per matrix, every O(N^3) flop on the FP64 tensor cores.
//
// Together with sb2st.cu this replaces the tridiagonalisation stage of Eigen::SelfAdjointEigenSolver
// as called from configuration_t::calc_ed (src/configuration.cpp:212-213).
//
// Look-ahead formulation: the trailing matrix is streamed through the SM ONCE per block column (one read,
// one write) instead of three times.  For block column k (columns k0..k0+7, trailing rows r0 = k0+8..N-1):
//   build   P_k = columns k0..k0+7 of A with the rank-16 update of block column k-1 applied on the fly
//           (thin, CUDA cores); the finished 8x8 diagonal block goes to band storage
//   QR      Householder QR of the m x 8 panel in shared memory -> V_k (unit lower trapezoidal), tau, R;
//           T_k (compact WY: Q = I - V T V^T) from the DMMA Gram matrix V^T V
//   export  V_k to global scratch in DMMA-fragment order (stays in L2: 320 KB per CTA)
//   pass    for every lower 32x32 tile of the trailing matrix, ONE visit:
//               A_tile -= V_{k-1} Z_{k-1}^T + Z_{k-1} V_{k-1}^T      (rank-16 SYR2K, DMMA)
//               store the tile
//               Y0_k(rows R) += A_tile V_k(rows C),  Y0_k(rows C) += A_tile^T V_k(rows R)   (SYMM, DMMA)
//   final   Y = Y0 T,  X = V^T Y (DMMA),  Z_k = Y - 1/2 V (T^T X);  export Z_k in fragment order
// The DMMA operands of the pass come from the fragment-ordered records in global memory (128-bit loads, L2 hits),
// which frees the shared memory the panels used to occupy for tile buffers: every warp owns TWO 8 KB tile
// buffers (16 tiles in flight per SM), filled by bulk asynchronous copies (cp.async.bulk + mbarrier transaction
// counts -- SASS UBLKCP) and drained by bulk stores, so the tensor-core warps never hold global loads of matrix
// data in registers and the next tile is in flight while the current one is in the tensor pipe.
// SYMM tile tasks are paired cyclically ({a, a-s}, s = 0..nt/2): a warp keeps the rows of its own tiles in
// registers for the whole pass and adds the partner rows straight into shared memory -- in one step all partner
// tiles are distinct, so no atomics are needed.
// R and the diagonal blocks are emitted in band storage AB[d][c] = A(c+d, c), d = 0..8.
//
// Matrix layout ("tiled"): only the lower-triangular 32x32 tiles are stored, tile (R, C), R >= C, at offset
// (R(R+1)/2 + C) * 1024 doubles, column-major inside the tile with an XOR swizzle of the row index
// (element (r, c) at c*32 + (r ^ 4*((c ^ c>>2) & 3))) -- unpadded, so a tile moves with ONE 8192-byte bulk copy in each
// direction, and all three fragment patterns (accumulator, direct operand, transposed operand) are bank-conflict-free.
// The tile grid is fixed in global coordinates; the panels V, Y, Z are indexed by global row and are zero above the
// trailing block, which makes the partial edge tiles of each block column come out right without masks.
// Requires N % 8 == 0 and N <= 1024.
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

// Developer ablation builds (make VARIANT=ablN EXTRA=-DFKMC_ABL=N; results are WRONG, only the timing is of interest):
//   1 no partner adds into shared memory   2 no transposed SYMM   3 operand records loaded once per pass   4 no tile stores
//   5 no tile loads (stale buffers)        6 no owned-row accumulation
#ifndef FKMC_ABL
#define FKMC_ABL 0
#endif

namespace {

constexpr int NB = 8;
constexpr int NW8 = 8;       // warps per CTA (N > 512); the N <= 512 variant runs 4 warps per CTA and two CTAs per SM
constexpr int TILE = 1024;   // doubles per tile (no padding)
constexpr int MAXOWN = 4;    // owned row tiles per warp: N <= 1024 -> nt <= 32 -> 4
constexpr int QR = 4;        // panel rows per thread during the QR: m <= 1016 -> 4

// column swizzle of tile row r: 4 * f(r & 7), f = (0, 2, 1, 3, 2, 0, 3, 1): rows {0,2,4,6} and {1,3,5,7} each get four different values
// (transposed operand reads of rows 2t + h are conflict-free) and rows 2j, 2j+1 differ in the high bit (so are the 128-bit
// accumulator-fragment accesses)
__host__ __device__ __forceinline__ int csw(int r) { return ((((r >> 1) & 3) + 2 * (r & 1)) & 3) << 2; }
// position of element (r, c) inside a tile (row-major, swizzled columns)
__host__ __device__ __forceinline__ int tix(int r, int c) { return (r << 5) + (c ^ csw(r)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// L2 eviction policies: the matrix tiles stream through L2 once per block column (evict first), the fragment records of the
// current block column are re-read by every tile task (evict last)
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes),
                 "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st_keep(double* p, double v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double2 ld_keep2(const double2* p, uint64_t pol) {
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Per-block progress counters in shared memory.  Writer: Y adds (all lanes), __syncwarp, counter store (lane 0); reader: counter load,
// then Y adds.  Release / acquire at CTA scope gives the happens-before edge the PTX memory model asks for; FKMC_S1_RELAXED builds the
// former volatile accesses (plain shared-memory accesses of one SM are performed in issue order by the LSU) for timing comparisons.
__device__ __forceinline__ int ld_flag(const int* p) {
    int v;
#ifdef FKMC_S1_RELAXED
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
#else
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
#endif
    return v;
}
__device__ __forceinline__ void st_flag(int* p, int v) {
#ifdef FKMC_S1_RELAXED
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
#else
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ void st_stream2(double* p, double a, double b, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
}

// DMMA without the volatile qualifier: a pure function of its operands, so ptxas may interleave independent accumulators
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__host__ __device__ __forceinline__ size_t tile_off(int R, int C) { return ((size_t)R * (R + 1) / 2 + C) * TILE; }
// element (i, j), i >= j, of the tiled matrix
__device__ __forceinline__ size_t elem_off(int i, int j) { return tile_off(i >> 5, j >> 5) + (size_t)tix(i & 31, j & 31); }

// Issue the bulk load of global tile (R, C) into a tile buffer (whole warp calls).
__device__ __forceinline__ void tile_load(double* buf, uint64_t* bar, const double* __restrict__ At, int R, int C, int lane, uint64_t pol) {
    __syncwarp();  // every lane is done with the previous contents
    if (lane == 0) {
        mbar_expect_tx(bar, (uint32_t)(TILE * 8));
        bulk_g2s(buf, At + tile_off(R, C), (uint32_t)(TILE * 8), bar, pol);
    }
}

// eight consecutive columns c0..c0+7 (c0 % 8 == 0) of row i of the tiled matrix, straight from L2 (four 128-bit loads)
__device__ __forceinline__ void load_row8(const double* __restrict__ At, int i, int c0, double (&out)[8]) {
    const int r = i & 31, cl = c0 & 31;
    const double* base = At + tile_off(i >> 5, c0 >> 5) + (r << 5);
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {
        const double2* q = reinterpret_cast<const double2*>(base + ((cl + 4 * hq) ^ csw(r)));
        const double2 a = __ldcg(q), b = __ldcg(q + 1);
        out[4 * hq] = a.x;
        out[4 * hq + 1] = a.y;
        out[4 * hq + 2] = b.x;
        out[4 * hq + 3] = b.y;
    }
}

// C(8x8) = P^T Q over rows [0, m): P, Q column-major [8][ld] (shared or global memory).  Each warp accumulates a
// slice of rows with DMMA, partial tiles are summed through red[nwarps][64]; result in out[a*8 + c].
__device__ __forceinline__ void gram8(const double* P, int ldp, const double* Q, int ldq, int m, double* red, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
    // A(g, k=t) = P[r+t][g], B(k=t, n=g) = Q[r+t][g]   (rows >= m are zero-padded); operands fetched eight row groups at a time
    for (int r0 = 4 * warp; r0 < m; r0 += 32 * nw) {
        double pa[8], qa[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + 4 * nw * i;
            pa[i] = (r < m) ? P[g * ldp + r + t] : 0.0;
            qa[i] = (r < m) ? Q[g * ldq + r + t] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            dmma(c0, c1, pa[i], qa[i]);
            dmma(d0, d1, pa[i + 1], qa[i + 1]);
        }
    }
    red[warp * 64 + g * 8 + 2 * t] = c0 + d0;
    red[warp * 64 + g * 8 + 2 * t + 1] = c1 + d1;
    __syncthreads();
    if (threadIdx.x < 64) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[w * 64 + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// one 64-byte fragment record (8 doubles per lane) from global memory
struct rec8 {
    double v[8];
};
__device__ __forceinline__ rec8 load_rec(const double* __restrict__ base, int R, int lane, uint64_t pol) {
    const double* p = base + ((size_t)(R * 32 + lane) << 3);
    rec8 r;
#pragma unroll
    for (int i = 0; i < 2; ++i)
        asm volatile("ld.global.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [%4], %5;"
                     : "=d"(r.v[4 * i]), "=d"(r.v[4 * i + 1]), "=d"(r.v[4 * i + 2]), "=d"(r.v[4 * i + 3])
                     : "l"(p + 4 * i), "l"(pol)
                     : "memory");
    return r;
}

// One visit of a staged tile S = (Rmax, Cmin):
//   S -= V_R Z_C^T + Z_R V_C^T                      (upd; records x[4q + blk] = X[32T + 8 blk + g][4q + t])
//   accR (rows of Rmax) += S V'[Cmin rows]          (accumulator fragments reused as A operands; vCp[2 blk + h] = V'[32C + 8 blk + 2t + h][g])
// Diagonal tiles hold both triangles, so this is all they need.
template <bool UPD>
__device__ __forceinline__ void tile_update_symm(double* __restrict__ Tb, double* __restrict__ gT, uint64_t pol, const rec8& uvR, const rec8& uzR,
                                                 const rec8& uvC, const rec8& uzC, const rec8& vCp, int lane, double (&accR)[4][2]) {
    const int g = lane >> 2, t = lane & 3;
    const int cs = csw(g);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        double acc[4][2][2];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int yy = 0; yy < 2; ++yy) {
                const double2 v = *reinterpret_cast<const double2*>(Tb + ((8 * x + g) << 5) + ((8 * (2 * half + yy) + 2 * t) ^ cs));
                acc[x][yy][0] = v.x;
                acc[x][yy][1] = v.y;
            }
        if (UPD) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int yy = 0; yy < 2; ++yy) dmma(acc[x][yy][0], acc[x][yy][1], -uvR.v[4 * q + x], uzC.v[4 * q + 2 * half + yy]);
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int yy = 0; yy < 2; ++yy) dmma(acc[x][yy][0], acc[x][yy][1], -uzR.v[4 * q + x], uvC.v[4 * q + 2 * half + yy]);
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int yy = 0; yy < 2; ++yy)
                {
                    const int off = ((8 * x + g) << 5) + ((8 * (2 * half + yy) + 2 * t) ^ cs);
                    *reinterpret_cast<double2*>(Tb + off) = make_double2(acc[x][yy][0], acc[x][yy][1]);  // for the transposed reads
                    if (FKMC_ABL != 4) st_stream2(gT + off, acc[x][yy][0], acc[x][yy][1], pol);                             // the tile itself, straight from registers
                }
        }
#pragma unroll
        for (int yy = 0; yy < 2; ++yy)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int x = 0; x < 4; ++x) dmma(accR[x][0], accR[x][1], acc[x][yy][h], vCp.v[2 * (2 * half + yy) + h]);
    }
}

// accC (rows of Cmin) += S^T V'[Rmax rows]   (off-diagonal tiles; same pair-order record: vRp[2 blk + h] = V'[32R + 8 blk + 2t + h][g])
__device__ __forceinline__ void tile_symm_t(const double* __restrict__ Tb, const rec8& vRp, int lane, double (&accC)[2][4][2]) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int rb = 0; rb < 4; ++rb)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = 8 * rb + 2 * t + h;
            const double* row = Tb + (r << 5);
            const int cs = csw(2 * t + h);
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) dmma(accC[h][cb][0], accC[h][cb][1], row[(8 * cb + g) ^ cs], vRp.v[2 * rb + h]);
        }
}

// The pass schedule.  Step s = 0..nt/2 pairs row block a with p = a - s (mod nt): tile (max, min).  Row blocks a < base = 8*(nt/8)
// are OWNED by warp a % 8 (their sums stay in its registers for the whole pass); the nt % 8 left-over blocks are FLOATING: their
// tasks rotate over the warps (at most one per warp and step, always last in the warp's step), and both halves of their result go
// to shared memory.  That keeps the tile count per warp and step equal to within one.
template <int NW>
struct sched {
    int warp, nt, smax, nown, base, rem;
    __device__ __forceinline__ int lim(int s) const { return (s > 0 && 2 * s == nt) ? (nt >> 1) : nt; }
    // row block of slot o in step s, or -1
    __device__ __forceinline__ int row(int s, int o) const {
        int a;
        if (o < nown) {
            a = warp + NW * o;
        } else if (o == nown) {
            const int k = (warp - s * rem) & (NW - 1);
            if (k >= rem) return -1;
            a = base + k;
        } else {
            return -1;
        }
        return a < lim(s) ? a : -1;
    }
    // advance (s, o) to this warp's next task; false when the pass is exhausted
    __device__ __forceinline__ bool next(int& s, int& o) const {
        ++o;
        while (s <= smax) {
            while (o <= nown) {
                if (row(s, o) >= 0) return true;
                ++o;
            }
            ++s;
            o = 0;
        }
        return false;
    }
    __device__ __forceinline__ void tile(int s, int o, int& a, int& Rmax, int& Cmin) const {
        a = row(s, o);
        int p = a - s;
        if (p < 0) p += nt;
        Rmax = max(a, p);
        Cmin = min(a, p);
    }
    __device__ __forceinline__ bool half_step(int s) const { return s > 0 && 2 * s == nt; }
    // number of shared-memory adds into row block b that precede its partner add / its floating own add of step s
    __device__ __forceinline__ int before_partner(int b, int s) const { return b >= base ? 2 * s - 1 : s - 1; }
    __device__ __forceinline__ int before_own(int b, int s) const {
        if (s == 0) return 0;
        const bool has_partner = !half_step(s) || b >= (nt >> 1);
        return 2 * s - 1 + (has_partner ? 1 : 0);
    }
};

// scratch layout per SM (doubles): Vcm [8][ld] | recVp [NT][32][8] | recAV [2][NT][32][8] | recAZ [NT][32][8]
__host__ __device__ __forceinline__ size_t scratch_doubles(int N) {
    const size_t mp = ((N + 31) / 32) * 32, nt = mp / 32;
    return 8 * (mp + 4) + 32 + 4 * nt * 256;
}

// DBG: cycle counters of CTA 0 (developer instrumentation, FKMC_S1_TIMING=1): dbg[0..7] phase totals of thread 0,
// dbg[8 + 8 w + i] pass sub-phase totals of warp w
#define S1_T(var) \
    long long var = 0; \
    if (DBG) var = clock64();
template <bool DBG, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 1 : 2)
sy2sb_kernel(double* __restrict__ A_all, size_t a_stride, int N, double* __restrict__ AB_all, double* __restrict__ scr_all, long long* dbg,
             unsigned* __restrict__ slot_mask) {
    long long ph_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ps_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, spin_t = 0;
    extern __shared__ __align__(128) double smem[];
    const int tid = threadIdx.x, T = NW * 32, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.x;
    double* At = A_all + (size_t)b * a_stride;  // tiled lower-triangular storage of this matrix
    double* AB = AB_all + (size_t)b * (NB + 1) * N;
    const int mp = ((N + 31) / 32) * 32, NT = mp >> 5;
    const int ld = mp + 4;    // V staging (X) and its global copy
    const int ldy = mp + 18;  // == 2 (mod 16): the accumulator-fragment adds into Y are bank-conflict-free
    const int xsz = max(8 * ld, NW * TILE);
    // shared memory: tile buffers 0 | X (V staging around the QR, tile buffers 1 during the pass) | Y | small stuff
    double* TB0 = smem;
    double* X = TB0 + NW * TILE;
    double* Y = X + xsz;
    double* G = Y + 8 * ldy;
    double* Tm = G + 64;
    double* M2 = Tm + 64;
    double* red = M2 + 64;                      // [NW*64] gram partials; QR: [2][NW][8] partials + [2][8] pivot row
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + NW * 64);
    int* blkstep = reinterpret_cast<int*>(bars + 2 * NW);  // [32] last step whose partner add into row block p is complete
    // global scratch
    // one scratch slot per SM (a single CTA fits on an SM): the records of the block column in flight stay hot in L2
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    // With two CTAs per SM (NW == 4) each takes one of the SM's two scratch slots: a bit per slot in slot_mask[smid], grabbed
    // here and released at the end of the kernel.
    __shared__ int my_slot;
    if (NW != 8) {
        if (tid == 0) {
            int got = -1;
            while (got < 0) {
                const unsigned old = atomicOr(slot_mask + smid, 1u);
                if (!(old & 1u)) got = 0;
                else {
                    const unsigned old2 = atomicOr(slot_mask + smid, 2u);
                    if (!(old2 & 2u)) got = 1;
                }
            }
            my_slot = got;
        }
        __syncthreads();
    }
    const int slot = (NW != 8) ? my_slot : 0;
    double* Vcm = scr_all + ((size_t)smid * (NW == 8 ? 1 : 2) + slot) * scratch_doubles(N);
    double* recVp = Vcm + 8 * ld + 32;
    double* recAV = recVp + NT * 256;
    double* recAZ = recAV + 2 * NT * 256;

    double* buf0 = TB0 + warp * TILE;
    double* buf1 = X + warp * TILE;
    uint64_t* bar0 = bars + 2 * warp;
    uint64_t* bar1 = bar0 + 1;
    uint32_t ph0 = 0, ph1 = 0;
    const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    int par = 0;

    for (int i = tid; i < 8 * ldy; i += T) Y[i] = 0.0;
    for (int i = tid; i < xsz; i += T) X[i] = 0.0;
    if (tid < 2 * NW) mbar_init(bars + tid, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // raw panel rows (loaded ahead of time: before the loop for the first panel, right after the pass for the others)
    double p[QR][NB], dg[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) dg[c] = 0.0;
    if (tid < NB) load_row8(At, tid, 0, dg);
#pragma unroll
    for (int q = 0; q < QR; ++q) {
        const int gi = NB + tid + T * q;
#pragma unroll
        for (int c = 0; c < NB; ++c) p[q][c] = 0.0;
        if (gi < N) load_row8(At, gi, 0, p[q]);
    }

    for (int k0 = 0; k0 < N; k0 += NB) {
        const int r0 = k0 + NB, m = N - r0;
        S1_T(t_a)
        // ---- build the panel in registers: columns k0..k0+7 of A, rows >= k0, minus the pending update of block column k-1.
        //      Thread tid owns local rows i = tid + 256 q (global row r0 + i); threads 0..7 also finish the diagonal block. ----
        if (k0 > 0 && tid < 64) {
            const int c = tid >> 3, j = tid & 7;  // rows k0+c of V_{k-1}, Z_{k-1}
            G[tid] = Vcm[j * ld + k0 + c];
            M2[tid] = Y[j * ldy + k0 + c];
        }
        __syncthreads();
        {
            const int gd = k0 + tid;  // diagonal-block row (tid < 8)
            if (k0 > 0) {
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    const int gi = r0 + tid + T * q;
                    if (gi < N) {
                        double vi[NB], zi[NB];
#pragma unroll
                        for (int j = 0; j < NB; ++j) {
                            vi[j] = Vcm[j * ld + gi];
                            zi[j] = Y[j * ldy + gi];
                        }
#pragma unroll
                        for (int c = 0; c < NB; ++c) {
                            double s = 0.0;
#pragma unroll
                            for (int j = 0; j < NB; ++j) s = fma(vi[j], M2[c * 8 + j], fma(zi[j], G[c * 8 + j], s));
                            p[q][c] -= s;
                        }
                    }
                }
                if (tid < NB) {
                    double vi[NB], zi[NB];
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        vi[j] = Vcm[j * ld + gd];
                        zi[j] = Y[j * ldy + gd];
                    }
#pragma unroll
                    for (int c = 0; c < NB; ++c) {
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < NB; ++j) s = fma(vi[j], M2[c * 8 + j], fma(zi[j], G[c * 8 + j], s));
                        dg[c] -= s;
                    }
                }
            }
            // finished diagonal block -> band storage
            if (tid < NB) {
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c <= tid) AB[(size_t)(tid - c) * N + k0 + c] = dg[c];
            }
        }
        if (m <= 0) break;
        S1_T(t_b)
        // ---- Householder QR of the panel, rows in registers, one block barrier per column ----
        double tauv[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            double* rbuf = red + (j & 1) * (NW * 8 + 8);
            tauv[j] = 0.0;
            if (m - j >= 2) {  // uniform
                // pw[c] = sum_{i>j} P[i][j] P[i][c], c >= j  (c = j gives the tail norm)
                double pw[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) pw[c] = 0.0;
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    const int i = tid + T * q;
                    if (i > j && i < m) {
#pragma unroll
                        for (int c = 0; c < NB; ++c)
                            if (c >= j) pw[c] = fma(p[q][j], p[q][c], pw[c]);
                    }
                }
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c >= j) {
                        const double s = warp_sum(pw[c]);
                        if (lane == 0) rbuf[warp * 8 + c] = s;
                    }
                if (tid == j) {
#pragma unroll
                    for (int c = 0; c < NB; ++c) rbuf[NW * 8 + c] = p[0][c];  // pivot row j: x0 and P[j][c]
                }
                __syncthreads();
                double Gs[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) {
                    Gs[c] = 0.0;
                    if (c >= j) {
#pragma unroll
                        for (int w = 0; w < NW; ++w) Gs[c] += rbuf[w * 8 + c];
                    }
                }
                const double tail2 = Gs[j];
                const double x0 = rbuf[NW * 8 + j];
                double beta = x0, inv = 0.0, tau = 0.0;
                if (tail2 > DBL_MIN) {
                    beta = sqrt(fma(x0, x0, tail2));
                    if (x0 >= 0.0) beta = -beta;
                    inv = 1.0 / (x0 - beta);
                    tau = (beta - x0) / beta;
                }
                tauv[j] = tau;
                // w_c = tau (P[j][c] + inv * G[c]);  rows i > j: v_i = P[i][j] inv, P[i][c] -= v_i w_c
                double w[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) w[c] = (c > j) ? tau * fma(inv, Gs[c], rbuf[NW * 8 + c]) : 0.0;
#pragma unroll
                for (int q = 0; q < QR; ++q) {
                    const int i = tid + T * q;
                    if (i > j && i < m) {
                        const double vi = p[q][j] * inv;
                        p[q][j] = vi;
#pragma unroll
                        for (int c = 0; c < NB; ++c)
                            if (c > j) p[q][c] = fma(-vi, w[c], p[q][c]);
                    }
                }
                if (tid == j) {  // row j itself (v_j = 1)
#pragma unroll
                    for (int c = 0; c < NB; ++c)
                        if (c > j) p[0][c] -= w[c];
                    p[0][j] = beta;
                }
            }
        }
        // R (upper triangle of the top 8x8) -> band storage; V explicit (unit lower trapezoidal) -> X staging
        if (tid < NB && tid < m) {
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (tid <= c) AB[(size_t)(NB + tid - c) * N + k0 + c] = p[0][c];
                if (tid < c) p[0][c] = 0.0;
                else if (tid == c) p[0][c] = 1.0;
            }
        }
#pragma unroll
        for (int q = 0; q < QR; ++q) {
            const int gi = r0 + tid + T * q;
            if (gi < ld) {
#pragma unroll
                for (int c = 0; c < NB; ++c) X[c * ld + gi] = (gi < N) ? p[q][c] : 0.0;
            }
        }
        // rows above the trailing block are zero in X: they were zeroed at start-up / by the previous export and only rows >= r0 are written
        __syncthreads();
        S1_T(t_c)
        // ---- T from the Gram matrix: T(0:j, j) = -tau_j T(0:j,0:j) G(0:j, j); lane i of warp 0 owns row i ----
        gram8(X + r0, ld, X + r0, ld, m, red, G);
        if (tid < NB) {
            double tr[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < NB; ++c)
                    if (c < j) s = (c >= tid) ? fma(tr[c], G[c * 8 + j], s) : s;
                tr[j] = (j == tid) ? tauv[j] : ((j > tid) ? -tauv[j] * s : 0.0);
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) Tm[tid * 8 + j] = tr[j];
        }
        // ---- export V: column-major copy + the fragment orders; clear Y for the new Y0; clear the X rows that die ----
        const int T0 = r0 >> 5, nt = NT - T0;  // tiles T0..NT-1 of the fixed global grid touch the trailing block
        for (int i = tid; i < 8 * ld; i += T) st_keep(Vcm + i, X[i], pol_keep);
        for (int i = tid; i < 8 * ldy; i += T) Y[i] = 0.0;
        {
            double* rA = recAV + (size_t)par * NT * 256;
            for (int idx = T0 * 256 + tid; idx < NT * 256; idx += T) {
                const int e = idx & 7, ln = (idx >> 3) & 31, R = idx >> 8, gg = ln >> 2, tt = ln & 3;
                // SYMM B operand (pair order): V[32R + 8 blk + 2t + h][g], e = 2 blk + h
                st_keep(recVp + idx, X[gg * ld + 32 * R + 8 * (e >> 1) + 2 * tt + (e & 1)], pol_keep);
                // SYR2K A/B operand: V[32R + 8 blk + g][4q + t], e = 4q + blk
                st_keep(rA + idx, X[(4 * (e >> 2) + tt) * ld + 32 * R + 8 * (e & 3) + gg], pol_keep);
            }
        }
        if (tid < 32) blkstep[tid] = 0;
        // generic-proxy accesses to X and the tile stores of the previous pass are ordered before the bulk copies of this pass
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
        S1_T(t_d)
        // ---- the pass: pending SYR2K of block column k-1 fused with the SYMM of block column k ----
        {
            const bool upd = k0 > 0;
            const double* uV = recAV + (size_t)(par ^ 1) * NT * 256;
            double own[MAXOWN][4][2];
#pragma unroll
            for (int o = 0; o < MAXOWN; ++o)
#pragma unroll
                for (int x = 0; x < 4; ++x) own[o][x][0] = own[o][x][1] = 0.0;
            sched<NW> sc;
            sc.warp = warp;
            sc.nt = nt;
            sc.smax = nt >> 1;
            sc.nown = nt / NW;
            sc.base = sc.nown * NW;
            sc.rem = nt - sc.base;
            // this warp's task list, packed (s | o << 5 | a << 8 | Rmax << 13 | Cmin << 18), task i in lane i % 32 of d[i / 32]
            int d0 = 0, d1 = 0, d2 = 0, ntasks = 0;
            for (int s = 0; s <= sc.smax; ++s)
                for (int o = 0; o <= sc.nown; ++o) {
                    const int a = sc.row(s, o);
                    if (a >= 0) {
                        int p = a - s;
                        if (p < 0) p += nt;
                        const int pk = s | (o << 5) | (a << 8) | (max(a, p) << 13) | (min(a, p) << 18);
                        if (ntasks == lane) d0 = pk;
                        if (ntasks == lane + 32) d1 = pk;
                        if (ntasks == lane + 64) d2 = pk;
                        ++ntasks;
                    }
                }
            auto task = [&](int i) { return __shfl_sync(0xffffffffu, i < 32 ? d0 : (i < 64 ? d1 : d2), i & 31); };
            // tile loads are issued two tasks ahead, operand records one task ahead
            uint32_t phb = ph0 | (ph1 << 1);
            rec8 uvR, uzR, uvC, uzC, vCp;
            if (ntasks > 0) {
                const int D = task(0), R = (D >> 13) & 31, C = (D >> 18) & 31;
                tile_load(buf0, bar0, At, T0 + R, T0 + C, lane, pol_stream);
                if (upd) {
                    uvR = load_rec(uV, T0 + R, lane, pol_keep);
                    uzR = load_rec(recAZ, T0 + R, lane, pol_keep);
                    uvC = load_rec(uV, T0 + C, lane, pol_keep);
                    uzC = load_rec(recAZ, T0 + C, lane, pol_keep);
                }
                vCp = load_rec(recVp, T0 + C, lane, pol_keep);
            }
            if (ntasks > 1) {
                const int D = task(1);
                tile_load(buf1, bar1, At, T0 + ((D >> 13) & 31), T0 + ((D >> 18) & 31), lane, pol_stream);
            }
            for (int i = 0; i < ntasks; ++i) {
                const int D = task(i);
                const int s = D & 31, o = (D >> 5) & 7, a = (D >> 8) & 31, Rmax = (D >> 13) & 31, Cmin = (D >> 18) & 31;
                const int slot = i & 1;
                const bool diag = (s == 0);
                double* Tb = slot ? buf1 : buf0;
                uint64_t* bar = slot ? bar1 : bar0;
                rec8 vRq;
                if (!diag) vRq = load_rec(recVp, T0 + Rmax, lane, pol_keep);
                S1_T(q0)
                if (FKMC_ABL != 5) {
                mbar_wait(bar, (phb >> slot) & 1);
                phb ^= 1u << slot;
                }
                S1_T(q1)
                double accR[4][2], accC[2][4][2];
#pragma unroll
                for (int x = 0; x < 4; ++x) accR[x][0] = accR[x][1] = accC[0][x][0] = accC[0][x][1] = accC[1][x][0] = accC[1][x][1] = 0.0;
                // entries outside the trailing block see a zero update (V, Z are zero there)
                if (upd) {
                    tile_update_symm<true>(Tb, At + tile_off(T0 + Rmax, T0 + Cmin), pol_stream, uvR, uzR, uvC, uzC, vCp, lane, accR);
                    __syncwarp();  // the transposed reads below see every lane's writes
                } else {
                    tile_update_symm<false>(Tb, nullptr, pol_stream, uvR, uzR, uvC, uzC, vCp, lane, accR);
                }
                S1_T(q2)
                // the operand records of the next task travel while this one finishes
                if (i + 1 < ntasks && FKMC_ABL != 3) {
                    const int D1 = task(i + 1), n1R = (D1 >> 13) & 31, n1C = (D1 >> 18) & 31;
                    if (upd) {
                        uvR = load_rec(uV, T0 + n1R, lane, pol_keep);
                        uzR = load_rec(recAZ, T0 + n1R, lane, pol_keep);
                        uvC = load_rec(uV, T0 + n1C, lane, pol_keep);
                        uzC = load_rec(recAZ, T0 + n1C, lane, pol_keep);
                    }
                    vCp = load_rec(recVp, T0 + n1C, lane, pol_keep);
                }
                S1_T(q3)
                if (!diag && FKMC_ABL != 2) tile_symm_t(Tb, vRq, lane, accC);
                S1_T(q4)
                // refill this buffer with the task after next once the store has read it
                {
                    if (i + 2 < ntasks && FKMC_ABL != 5) {
                        const int D2 = task(i + 2);
                        tile_load(Tb, bar, At, T0 + ((D2 >> 13) & 31), T0 + ((D2 >> 18) & 31), lane, pol_stream);
                    }
                }
                S1_T(q5)
                // owned rows stay in registers; everything else goes to shared memory.  The adds into one row block happen in a fixed
                // order (per step: partner add, then the floating own add), enforced with a per-block counter: deterministic sums.
                const bool fl = (o == sc.nown);
                double ow[4][2];
                if (diag) {
#pragma unroll
                    for (int x = 0; x < 4; ++x) { ow[x][0] = accR[x][0]; ow[x][1] = accR[x][1]; }
                } else {
                    const bool ownR = (a == Rmax);
                    const int pblk = ownR ? Cmin : Rmax;
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const double c0 = accC[0][x][0] + accC[1][x][0], c1 = accC[0][x][1] + accC[1][x][1];
                        ow[x][0] = ownR ? accR[x][0] : c0;
                        ow[x][1] = ownR ? accR[x][1] : c1;
                        accR[x][0] = ownR ? c0 : accR[x][0];
                        accR[x][1] = ownR ? c1 : accR[x][1];
                    }
                    const int need = sc.before_partner(pblk, s);
                    S1_T(sp0)
                    if (FKMC_ABL != 1) {
                    while (ld_flag(blkstep + pblk) < need) {
                    }
                    if (DBG) spin_t += clock64() - sp0;
                    double* yp = Y + 32 * (T0 + pblk) + g + 2 * t * ldy;
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int h = 0; h < 2; ++h) yp[h * ldy + 8 * x] += accR[x][h];
                    __syncwarp();
                    if (lane == 0) st_flag(blkstep + pblk, need + 1);
                    } else {
                        // keep the values alive so that the DMMAs that produced them are not optimised away
                        double sink = 0.0;
#pragma unroll
                        for (int x = 0; x < 4; ++x) sink += accR[x][0] + accR[x][1];
                        if (sink == 1.2345e300) Y[lane] = sink;
                    }
                }
                if (!fl && FKMC_ABL == 6) {
                    double sink = 0.0;
#pragma unroll
                    for (int x = 0; x < 4; ++x) sink += ow[x][0] + ow[x][1];
                    if (sink == 1.2345e300) own[0][0][0] = sink;
                } else if (!fl) {
#pragma unroll
                    for (int oo = 0; oo < MAXOWN; ++oo)
                        if (oo == o) {
#pragma unroll
                            for (int x = 0; x < 4; ++x) { own[oo][x][0] += ow[x][0]; own[oo][x][1] += ow[x][1]; }
                        }
                } else {
                    const int need2 = sc.before_own(a, s);
                    while (FKMC_ABL != 1 && ld_flag(blkstep + a) < need2) {
                    }
                    double* yo = Y + 32 * (T0 + a) + g + 2 * t * ldy;
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int h = 0; h < 2; ++h) yo[h * ldy + 8 * x] += ow[x][h];
                    __syncwarp();
                    if (lane == 0) st_flag(blkstep + a, need2 + 1);
                }
                if (DBG) {
                    const long long q6 = clock64();
                    ps_t[0] += q1 - q0;  // wait for the tile
                    ps_t[1] += q2 - q1;  // update + accR + store issue
                    ps_t[2] += q3 - q2;  // next-record issue
                    ps_t[3] += q4 - q3;  // transposed SYMM
                    ps_t[4] += q5 - q4;  // store-read wait + refill
                    ps_t[5] += q6 - q5;  // adds
                    ps_t[6] += 1;
                }
            }
            ph0 = phb & 1;
            ph1 = (phb >> 1) & 1;
            S1_T(t_e0)
            __syncthreads();  // every warp's partner adds are in
            S1_T(t_e)
            if (DBG) {
                ph_t[0] += t_b - t_a;
                ph_t[1] += t_c - t_b;
                ph_t[2] += t_d - t_c;
                ph_t[3] += t_e - t_d;
                ph_t[5] += t_e - t_e0;
                ph_t[6] = t_e;
            }
#pragma unroll
            for (int o = 0; o < MAXOWN; ++o) {
                const int a = warp + NW * o;
                if (o < nt / NW) {
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int h = 0; h < 2; ++h) Y[(2 * t + h) * ldy + 32 * (T0 + a) + 8 * x + g] += own[o][x][h];
                }
            }
        }
        // ---- raw rows of the next panel (columns r0..r0+7): in flight while Z is finished ----
        {
#pragma unroll
            for (int c = 0; c < NB; ++c) dg[c] = 0.0;
            if (tid < NB) load_row8(At, r0 + tid, r0, dg);
#pragma unroll
            for (int q = 0; q < QR; ++q) {
                const int gi = r0 + NB + tid + T * q;
#pragma unroll
                for (int c = 0; c < NB; ++c) p[q][c] = 0.0;
                if (gi < N) load_row8(At, gi, r0, p[q]);
            }
        }
        __syncthreads();
        // ---- Y = Y0 T;  X = V^T Y;  Z = Y - 1/2 V (T^T X)   (V from its global copy: the staging buffer held tiles).
        //      Rows of the first (edge) tile above the trailing block picked up band entries: Z must be zero there. ----
        double* YV = Y + r0;
        for (int idx = tid; idx < 8 * (r0 - 32 * T0); idx += T) {
            const int j = idx / (r0 - 32 * T0), gi = 32 * T0 + idx % (r0 - 32 * T0);
            Y[j * ldy + gi] = 0.0;
        }
        for (int i = tid; i < m; i += T) {
            double y0[NB], y[NB];
#pragma unroll
            for (int a = 0; a < NB; ++a) y0[a] = YV[a * ldy + i];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int a = 0; a < NB; ++a)
                    if (a <= c) sacc = fma(y0[a], Tm[a * 8 + c], sacc);
                y[c] = sacc;
            }
#pragma unroll
            for (int c = 0; c < NB; ++c) YV[c * ldy + i] = y[c];
        }
        __syncthreads();
        gram8(Vcm + r0, ld, YV, ldy, m, red, G);  // G = X = V^T Y
        if (tid < 64) {
            const int a = tid >> 3, c = tid & 7;  // M2 = T^T X
            double sacc = 0.0;
            for (int q = 0; q <= a; ++q) sacc += Tm[q * 8 + a] * G[q * 8 + c];
            M2[a * 8 + c] = sacc;
        }
        __syncthreads();
        for (int i = tid; i < m; i += T) {
            double v[NB];
#pragma unroll
            for (int a = 0; a < NB; ++a) v[a] = Vcm[a * ld + r0 + i];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double sacc = 0.0;
#pragma unroll
                for (int a = 0; a < NB; ++a) sacc = fma(v[a], M2[a * 8 + c], sacc);
                YV[c * ldy + i] -= 0.5 * sacc;
            }
        }
        __syncthreads();
        // export Z in SYR2K fragment order; X rows that leave the trailing block must read as zero from now on
        for (int idx = T0 * 256 + tid; idx < NT * 256; idx += T) {
            const int e = idx & 7, ln = (idx >> 3) & 31, R = idx >> 8, gg = ln >> 2, tt = ln & 3;
            st_keep(recAZ + idx, Y[(4 * (e >> 2) + tt) * ldy + 32 * R + 8 * (e & 3) + gg], pol_keep);
        }
        for (int i = tid; i < xsz; i += T) X[i] = 0.0;
        par ^= 1;
        if (DBG) ph_t[4] += clock64() - ph_t[6], ph_t[6] = 0;
        // (the __syncthreads after the staging of G/M2 at the top of the next iteration orders these writes)
    }
    if (NW != 8) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAnd(slot_mask + smid, ~(1u << slot));
        }
    }
    if (DBG && blockIdx.x == 0) {
        if (tid == 0)
            for (int i = 0; i < 8; ++i) dbg[i] = ph_t[i];
        if (lane == 0)
            for (int i = 0; i < 8; ++i) dbg[8 + 8 * warp + i] = (i == 2) ? ps_t[i] : ps_t[i];
        if (lane == 0) dbg[72 + warp] = spin_t;
    }
}

}  // namespace

// warps per CTA: 8 (one CTA per SM) above N = 512, else 4 with two CTAs per SM, so that the panel phases of one matrix run
// beside the tensor-core pass of another (the QR keeps four panel rows per thread: m <= 4 * 128)
static int sy2sb_warps(int N) {
    if (const char* e = getenv("FKMC_S1_WARPS")) return atoi(e) == 4 && N <= 512 ? 4 : 8;  // (environment: developer override)
    return N > 512 ? 8 : 4;
}

size_t fkmc_sy2sb_smem(int N) {
    const size_t nw = sy2sb_warps(N);
    const size_t mp = ((N + 31) / 32) * 32, ld = mp + 4;
    const size_t xsz = std::max<size_t>(8 * ld, nw * TILE);
    return sizeof(double) * (nw * TILE + xsz + 8 * (mp + 18) + 3 * 64 + nw * 64) + sizeof(uint64_t) * 2 * nw + sizeof(int) * 32 + 16;
}
size_t fkmc_sy2sb_scratch(int N) { return scratch_doubles(N); }

int fkmc_launch_sy2sb_small(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_AB);

__global__ void nsmid_kernel(unsigned* out) {
    unsigned n;
    asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n));
    *out = n;
}

// doubles per matrix in the tiled lower-triangular layout
size_t fkmc_tiled_stride(int N) {
    const size_t nt = (N + 31) / 32;
    return nt * (nt + 1) / 2 * TILE;
}
bool fkmc_use_tiled(const fkmc_ctx* ctx, int N) { return N >= ctx->tiled_min && N % 8 == 0 && N <= 1024; }

// dense -> band on a batch of matrices in the tiled layout (see fkmc_launch_build_h_tiled / fkmc_launch_to_tiled)
int fkmc_launch_sy2sb_tiled(fkmc_ctx* ctx, double* d_At, int N, int B, double* d_AB) {
    fkmc_prof_scope ps(ctx, "sy2sb");
    if (!fkmc_use_tiled(ctx, N)) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sy2sb_tiled: needs 512 <= N <= 1024, N % 8 == 0");
    const int nw = sy2sb_warps(N), slots = nw == 8 ? 1 : 2;
    const size_t smem = fkmc_sy2sb_smem(N);
    if (smem > ctx->smem_optin) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sy2sb: matrix too large for shared memory");
    // fragment-ordered panel records, one slot per SM
    if (ctx->nsmid == 0) {
        unsigned* d_n = nullptr;
        unsigned h_n = 0;
        FKMC_CUDA(ctx, cudaMalloc(&d_n, sizeof(unsigned)));
        nsmid_kernel<<<1, 1, 0, ctx->stream>>>(d_n);
        FKMC_CUDA(ctx, cudaMemcpyAsync(&h_n, d_n, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d_n);
        ctx->nsmid = (int)h_n;
    }
    const size_t need = fkmc_sy2sb_scratch(N) * (size_t)ctx->nsmid * slots + (size_t)ctx->nsmid;  // + one word per SM: the slot bits
    if (need > ctx->s1_scratch_cap) {
        if (ctx->d_s1_scratch) {
            FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_s1_scratch);
            ctx->d_s1_scratch = nullptr;
            ctx->s1_scratch_cap = 0;
        }
        FKMC_CUDA(ctx, cudaMalloc(&ctx->d_s1_scratch, sizeof(double) * need));
        ctx->s1_scratch_cap = need;
        FKMC_CUDA(ctx, cudaMemsetAsync(ctx->d_s1_scratch, 0, sizeof(double) * need, ctx->stream));
    }
    if (getenv("FKMC_S1_TIMING")) {
        long long* d_dbg = nullptr;
        long long h[80];
        FKMC_CUDA(ctx, cudaMalloc(&d_dbg, sizeof(h)));
        FKMC_CUDA(ctx, cudaMemsetAsync(d_dbg, 0, sizeof(h), ctx->stream));
        if (nw != 8) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "FKMC_S1_TIMING needs N > 512");
        FKMC_CUDA(ctx, cudaFuncSetAttribute(sy2sb_kernel<true, NW8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sy2sb_kernel<true, NW8><<<B, NW8 * 32, smem, ctx->stream>>>(d_At, fkmc_tiled_stride(N), N, d_AB, ctx->d_s1_scratch, d_dbg, nullptr);
        FKMC_CUDA(ctx, cudaMemcpyAsync(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(d_dbg);
        fprintf(stderr, "[s1 timing, CTA 0, cycles] build %lld  qr %lld  T+export %lld  pass %lld (end barrier %lld)  final %lld\n", h[0], h[1], h[2],
                h[3], h[5], h[4]);
        for (int w = 0; w < NW8; ++w)
            fprintf(stderr, "  warp %d: tasks %lld | tile wait %lld  update %lld  rec issue %lld  symm_t %lld  refill %lld (store-read wait %lld)  adds %lld (spin %lld)\n", w,
                    h[8 + 8 * w + 6], h[8 + 8 * w], h[8 + 8 * w + 1], h[8 + 8 * w + 2], h[8 + 8 * w + 3], h[8 + 8 * w + 4], h[8 + 8 * w + 7],
                    h[8 + 8 * w + 5], h[72 + w]);
        ctx->launches++;
        return FKMC_OK;
    }
    // the slot bits live behind the scratch slots (zeroed with the scratch; every CTA clears its bit when it ends)
    unsigned* slot_mask = reinterpret_cast<unsigned*>(ctx->d_s1_scratch + fkmc_sy2sb_scratch(N) * (size_t)ctx->nsmid * slots);
    if (nw == 8) {
        FKMC_CUDA(ctx, cudaFuncSetAttribute(sy2sb_kernel<false, NW8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sy2sb_kernel<false, NW8><<<B, NW8 * 32, smem, ctx->stream>>>(d_At, fkmc_tiled_stride(N), N, d_AB, ctx->d_s1_scratch, nullptr, slot_mask);
    } else {
        FKMC_CUDA(ctx, cudaFuncSetAttribute(sy2sb_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FKMC_CUDA(ctx, cudaMemsetAsync(slot_mask, 0, sizeof(unsigned) * (size_t)ctx->nsmid, ctx->stream));
        sy2sb_kernel<false, 4><<<B, 4 * 32, smem, ctx->stream>>>(d_At, fkmc_tiled_stride(N), N, d_AB, ctx->d_s1_scratch, nullptr, slot_mask);
    }
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

__device__ __forceinline__ void tile_decode(int tile, int& R, int& C) {
    R = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
    while (R * (R + 1) / 2 > tile) --R;
    while ((R + 1) * (R + 2) / 2 <= tile) ++R;
    C = tile - R * (R + 1) / 2;
}

// column-major [B][N][N] (lower triangle) -> tiled layout
__global__ void __launch_bounds__(256) to_tiled_kernel(const double* __restrict__ A_all, int N, double* __restrict__ At_all, size_t t_stride) {
    const int b = blockIdx.y, tile = blockIdx.x;
    int R, C;
    tile_decode(tile, R, C);
    const double* A = A_all + (size_t)b * N * N;
    double* dst = At_all + (size_t)b * t_stride + (size_t)tile * TILE;
    for (int e = threadIdx.x; e < TILE; e += blockDim.x) {
        const int r = e >> 5, c = (e & 31) ^ csw(r);
        const int i = 32 * R + r, j = 32 * C + c;
        double v = 0.0;
        if (i < N && j < N) v = A[(size_t)min(i, j) * N + max(i, j)];  // symmetric fill from the lower triangle
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
                                                            double* __restrict__ At_all, size_t t_stride, const unsigned char* __restrict__ tile_mask) {
    const int b = blockIdx.y, tile = blockIdx.x;
    double* dst = At_all + (size_t)b * t_stride + (size_t)tile * TILE;
    if (!tile_mask[tile]) {
        // no hopping and no diagonal entry falls into this tile (most tiles of a lattice Hamiltonian): plain 16-byte zero stores
        double2* d2 = reinterpret_cast<double2*>(dst);
        for (int e = threadIdx.x; e < TILE / 2; e += blockDim.x) d2[e] = make_double2(0.0, 0.0);
        return;
    }
    int R, C;
    tile_decode(tile, R, C);
    const int32_t* fb = f + (size_t)b * N;
    for (int e = threadIdx.x; e < TILE; e += blockDim.x) {
        const int r = e >> 5, c = (e & 31) ^ csw(r);
        const int i = 32 * R + r, j = 32 * C + c;
        double v = 0.0;
        if (i < N && j < N) {
            if (i == j) {
                v = U * (double)fb[i] - mu_c;
            } else {
                // the lower triangle is the live one (as in the reference); diagonal tiles carry its mirror image
                const int lo = max(i, j), up = min(i, j);
                for (int z = 0; z < Z; ++z)
                    if (nbr_idx[z * N + lo] == up) v = nbr_val[z * N + lo];
            }
        }
        dst[e] = v;
    }
}

int fkmc_launch_build_h_tiled(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double* d_At) {
    fkmc_prof_scope ps(ctx, "build_h");
    const int N = ctx->N, nt = (N + 31) / 32, ntiles = nt * (nt + 1) / 2;
    if (!ctx->d_tile_mask) {
        // tiles that can hold a non-zero: the diagonal ones and those a stencil entry of the lower triangle falls into
        std::vector<unsigned char> mask(ntiles, 0);
        for (int R = 0; R < nt; ++R) mask[(size_t)R * (R + 1) / 2 + R] = 1;
        for (int z = 0; z < ctx->Z; ++z)
            for (int i = 0; i < N; ++i) {
                const int j = ctx->h_nbr_idx[(size_t)z * N + i];
                if (j >= N) continue;
                const int lo = std::max(i, j), up = std::min(i, j);
                mask[(size_t)(lo >> 5) * ((lo >> 5) + 1) / 2 + (up >> 5)] = 1;
            }
        FKMC_CUDA(ctx, cudaMalloc(&ctx->d_tile_mask, ntiles));
        FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_tile_mask, mask.data(), ntiles, cudaMemcpyHostToDevice, ctx->stream));
        FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    dim3 grid(ntiles, B);
    build_h_tiled_kernel<<<grid, 256, 0, ctx->stream>>>(d_f, ctx->d_nbr_idx, ctx->d_nbr_val, N, ctx->Z, U, mu_c, d_At, fkmc_tiled_stride(N),
                                                       ctx->d_tile_mask);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

// column-major entry point (stage-level API and small matrices)
int fkmc_launch_sy2sb(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_AB) {
    if (!fkmc_use_tiled(ctx, N)) return fkmc_launch_sy2sb_small(ctx, d_A, N, B, d_AB);
    // convert in place is not possible (layouts overlap): stage through a scratch allocation
    double* d_At = nullptr;
    FKMC_CUDA(ctx, cudaMalloc(&d_At, sizeof(double) * fkmc_tiled_stride(N) * (size_t)B));
    int rc = fkmc_launch_to_tiled(ctx, d_A, N, B, d_At);
    if (!rc) rc = fkmc_launch_sy2sb_tiled(ctx, d_At, N, B, d_AB);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_At);
    return rc;
}
