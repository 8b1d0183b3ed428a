// Band -> band reduction for lattice Hamiltonians (test infrastructure: none; this is product code).
//
// The reference diagonalises the dense N x N matrix of configuration_t::calc_hamiltonian (configuration.cpp:60-67) with Eigen's
// SelfAdjointEigenSolver (configuration.cpp:76-105).  For the two-dimensional lattices that matrix is sparse: after folding the slow
// coordinate (y -> 0, L-1, 1, L-2, ...) the periodic lattice becomes a band matrix of half-bandwidth 2L (2L+1 triangular), 64 for the
// 32 x 32 lattice of the headline configuration.  Eigenvalues do not depend on the ordering of the sites, so the eigenvalue-only path
// can start from the band: this file reduces half-bandwidth <= 64 to half-bandwidth 8 by block bulge chasing (Bischof-Lang-Sun SBR
// scheme, QR of a 64 x 8 panel, compact-WY two-sided updates of 64 x 64 blocks, all O(b N^2) flops as FP64 DMMA), 6*56*N^2 flop
// instead of the 4/3 N^3 of the dense reduction (3.6x fewer at N = 1024), and hands the result to sb2st.cu / tridiag_eig.cu.
//
// Storage: 8 x 8 tiles, tile (I, J) of the lower band with 0 <= I - J <= 15 (band + bulge) at band[(I*16 + J - I + 15)*64], each tile
// row-major = exactly the DMMA accumulator layout (lane t holds the double2 at 16 t bytes): tile loads and stores are one fully
// coalesced 512-byte access per warp.  Rows N .. N+127 are zero padding: every step has the same shape.
//
// One CTA (8 warps) per matrix, two CTAs per SM.  Sweep j removes columns 8j..8j+7 below the 8-band; step p of the sweep works on rows
// R_p = [8j + 8 + 64p, +64):  QR of the panel (warp 0, shuffles), then with Q = I - V T V^T
//     B = A[R_p, R_(p-1)]  <- Q^T B          (row-owned tiles; W = T^T V^T B computed column-owned)
//     S = A[R_p, R_p]      <- Q^T S Q        (X = S V T,  M = T^T V^T X,  Y = X - V M / 2,  S -= V Y^T + Y V^T)
//     G = A[R_(p+1), R_p]  <- G Q            (entirely inside the owning warp; G stays in shared memory as the next step's panel + B)
// Each warp owns one 8-row block of S, G and B in registers in accumulator layout; an accumulator tile is used directly as the A operand
// of the next DMMA by permuting the contraction index (column 2q+kk of the tile is contraction slot q of k-step kk), the matching B operands
// come from row- or column-permuted copies of V in shared memory (all fragment loads bank-conflict free: row stride 12).
#include "common.cuh"

namespace {

constexpr int VS = 12;                 // row stride (doubles) of the 64 x 8 operand arrays in shared memory
constexpr int NT = 16;                 // tile diagonals stored per block row
// FP64 arbitration (tools/ubench/contend.cu): a warp issuing scalar DFMA next to DMMA-streaming warps of the same sub-partition waits 73
// cycles per instruction with one such neighbour and starves with two; warps of other sub-partitions do not matter.  The panel
// factorisation is a serial chain of scalar FP64 work, so it gets a sub-partition without tensor work: the CTA has eleven warps, those
// whose hardware slot (%warpid) is a multiple of four hold the panel warp (the other one or two idle), the eight compute warps are
// taken from the rest.  Two CTAs per SM: both panel warps share sub-partition 0, the 16 compute warps the other three.
// (Measured at N = 1024, 1024 matrices: panel warp next to compute warps 31.9 ms, 13.2k cycles per factorisation; isolated 6.4k.)
#ifndef FKMC_SBR_WARPS
#define FKMC_SBR_WARPS 10
#endif
constexpr int SBR_WARPS = FKMC_SBR_WARPS;
constexpr int SBR_THREADS = SBR_WARPS * 32;
constexpr int SBR_CTAS = 2;
constexpr int SBR_SYNC = 288;          // threads on the panel / compute named barriers (panel warp + eight compute warps)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// named barriers: 1 = reflectors ready (panel warp arrives, compute warps wait), 2 = next panel ready (compute warps arrive, panel warp
// waits), 3 = compute warps only
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// The sweep barrier of the whole CTA.  Every role (panel warp, compute warps, idle warps) calls it from its own loop; keeping it in one
// non-inlined function makes that a single barrier instruction (the tools that check barrier convergence key on the instruction).
__device__ __noinline__ void sweep_barrier() { __syncthreads(); }
__device__ __forceinline__ void bar_arrive(int id, int n) {
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// word offset of element (row, col) of a swizzled 8 x 8 tile in shared memory (row pairs 2,3 and 6,7 have their column halves exchanged:
// accumulator-layout double2 accesses stay contiguous, (row = 4kk + q, col = g) operand reads hit 16 distinct banks per half warp)
__device__ __forceinline__ int swz(int row, int col) { return row * 8 + (col ^ (((row >> 1) & 1) << 2)); }

struct vset {
    double Vn[64 * VS];     // V[row][refl]
    double Vp[64 * VS];     // rows permuted inside each tile: row r at (r & 1)*4 + (r >> 1)
    double Vq[64 * VS];     // reflector index permuted the same way
    double Tn[8 * VS];      // T[row][col]
};
struct sbr_smem {
    double G[4096];         // the B block of the current step = G of the previous one: tile-major (tile (k, c) at (k*8 + c)*64), swizzled tiles
    double P[512];          // the next panel (first tile column of the updated G), tile k at k*64, swizzled
    vset V[2];              // double-buffered: the panel warp factors step p+1 while the compute warps apply step p
    double XY[64 * VS];     // X = S V T (rows read back by the owning warp only), later overwritten by Y
    double W[64 * VS];      // W[col][refl] = (T^T V^T B)^T
    double Zp[8][64];       // per-warp partials of (V^T X)^T
    double pad[368];   // the panel warp's reduction scratch
};
static_assert(sizeof(sbr_smem) <= 115712, "two CTAs per SM");

__device__ __forceinline__ double* tile_ptr(double* band, int I, int J) { return band + ((size_t)I * NT + (J - I + NT - 1)) * 64; }

// reciprocal square root / reciprocal: MUFU seed + Newton steps
__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double rcp_seed(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// Warp-wide sums of eight per-lane partials plus the broadcast of one lane's eight values, through shared memory (shuffles and shared
// memory use the same queue, which the compute warps keep busy: fewer, wider operations win).  The 32 x 8 partials are transposed through
// the pad (row stride 10 doubles, each group of eight rows shifted by a further 8: 128-bit stores and 64-bit column loads conflict free), lane l adds
// the eight partials of column l % 8 held by its group of eight lanes, two shuffle stages add the four groups, the totals and the
// source lane's row go back through the pad.  pad: 368 doubles owned by the warp, 16-byte aligned.
__device__ __forceinline__ void warp_sum8_bcast(double (&v)[8], double (&h)[8], int src, int lane, double* pad) {
    double* row = pad + lane * 10 + (lane >> 3) * 8;
#pragma unroll
    for (int k = 0; k < 8; k += 2) *reinterpret_cast<double2*>(row + k) = make_double2(v[k], v[k + 1]);
    double* hrow = pad + 352;
    if (lane == src) {
#pragma unroll
        for (int k = 0; k < 8; k += 2) *reinterpret_cast<double2*>(hrow + k) = make_double2(h[k], h[k + 1]);
    }
    __syncwarp();
    const int gb = lane & 24;
    const double* col = pad + gb * 10 + (gb >> 3) * 8 + (lane & 7);
    double x = ((col[0] + col[10]) + (col[20] + col[30])) + ((col[40] + col[50]) + (col[60] + col[70]));
    x += __shfl_xor_sync(0xffffffffu, x, 8);
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    double* tot = pad + 360;
    if (lane < 8) tot[lane] = x;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const double2 a = *reinterpret_cast<const double2*>(tot + k), b = *reinterpret_cast<const double2*>(hrow + k);
        v[k] = a.x; v[k + 1] = a.y; h[k] = b.x; h[k + 1] = b.y;
    }
    __syncwarp();   // the pad is rewritten by the next call
}

// Householder QR of a 64 x 8 panel held by one warp (lane l: rows l and l + 32 in a0 / a1).  Branch-free; one reduction round per
// reflector: the raw products x_i^T a_k over the rows below the diagonal give the norm (k = i), the projections v^T a_k (k > i) and, with
// the earlier reflectors in the slots k < i, column i of V^T V for the compact-WY factor T (row l of T lives in lane l).
// Writes V (three layouts) and T to vs, and the factored panel ([R; 0]) to the global tiles (I0 + k, Jp), k = 0..7.
#ifdef FKMC_SBR_TIMING
__device__ __forceinline__ void qr_panel(double (&a0)[8], double (&a1)[8], vset& vs, double* pad, double* band, int I0, int Jp, int lane, bool arrive, long long* qt) {
#else
__device__ __forceinline__ void qr_panel(double (&a0)[8], double (&a1)[8], vset& vs, double* pad, double* band, int I0, int Jp, int lane, bool arrive) {
#endif
#ifdef FKMC_SBR_TIMING
    long long q0 = clock64();
#define QR_TICK(i) { const long long n__ = clock64(); qt[i] += n__ - q0; q0 = n__; }
#else
#define QR_TICK(i)
#endif
    const int r0 = lane, r1 = lane + 32, l = lane & 7;
    const int p0 = (r0 & ~7) + ((r0 & 1) * 4 + ((r0 & 7) >> 1)), p1 = (r1 & ~7) + ((r1 & 1) * 4 + ((r1 & 7) >> 1));
    double trow[8], rdiag[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // columns k < i of the register panel hold the earlier reflectors below their diagonal (LAPACK storage), so one formula serves the
        // projections (k > i), the norm (k = i) and column i of V^T V (k < i)
        const double m0 = (lane > i) ? a0[i] : 0.0;
        double r[8], head[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            r[k] = fma(m0, a0[k], a1[i] * a1[k]);
            head[k] = a0[k];   // row i (lane i) is broadcast together with the sums
        }
        QR_TICK(0)
        warp_sum8_bcast(r, head, i, lane, pad);
        QR_TICK(1)
        const double sigma = r[i], alpha = head[i];
        const bool zero = sigma == 0.0;
        const double n2 = zero ? 1.0 : fma(alpha, alpha, sigma), aa = fabs(alpha);
        double rn = rsqrt_seed(n2);
        double rc = rcp_seed(fma(n2, rn, aa));   // seed of 1 / (|alpha| + norm), refined below against the accurate norm
        {   // seed good to 2^-20: one third-order step reaches rounding level (tools/ubench/seedacc.cu)
            const double e = fma(-n2 * rn, rn, 1.0);
            rn = fma(rn * e, fma(e, 0.375, 0.5), rn);
        }
        const double nrm = n2 * rn, d = aa + nrm;
#pragma unroll
        for (int it = 0; it < 2; ++it) rc = fma(rc, fma(-d, rc, 1.0), rc);
        const double s = zero ? 0.0 : copysign(rc, alpha);             // 1 / (alpha - beta), beta = -sign(alpha) norm
        const double tau = zero ? 0.0 : fma(aa, rn, 1.0);              // (beta - alpha) / beta
        QR_TICK(2)
        const double v0 = (lane > i) ? a0[i] * s : ((lane == i) ? 1.0 : 0.0), v1 = a1[i] * s;
#pragma unroll
        for (int k = i + 1; k < 8; ++k) {
            const double wk = tau * fma(s, r[k], head[k]);              // tau v^T a_k
            a0[k] = fma(-wk, v0, a0[k]);
            a1[k] = fma(-wk, v1, a1[k]);
        }
        rdiag[i] = zero ? alpha : -copysign(nrm, alpha);
        // column i of T: T[l][i] = -tau sum_{m=l}^{i-1} T[l][m] (V_m^T v_i),  V_m^T v_i = V_m[i] + s * raw_m
        double t = 0.0;
#pragma unroll
        for (int m = 0; m < i; ++m) t = fma(trow[m], fma(s, r[m], head[m]), t);   // trow[m] = 0 for m < l
        trow[i] = (l == i) ? tau : ((l < i) ? -tau * t : 0.0);
        // v_i replaces the column below the diagonal (lane i keeps the implicit 1 as 0 products: row i is never part of a later sum)
        if (lane >= i) a0[i] = v0;
        a1[i] = v1;
        const int pi = (i & 1) * 4 + (i >> 1);
        vs.Vn[r0 * VS + i] = v0;
        vs.Vn[r1 * VS + i] = v1;
        vs.Vp[p0 * VS + i] = v0;
        vs.Vp[p1 * VS + i] = v1;
        vs.Vq[r0 * VS + pi] = v0;
        vs.Vq[r1 * VS + pi] = v1;
        QR_TICK(3)
    }
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) vs.Tn[l * VS + i] = trow[i];
    }
    if (arrive) bar_arrive(1, SBR_SYNC);   // V and T are complete: the compute warps start while the factored panel goes out
    // factored panel to global memory: R in the first tile (rows 0..7 = lanes 0..7), zeros below
    {
        const int k0 = lane >> 3, r = lane & 7;
        double* t0 = tile_ptr(band, I0 + k0, Jp) + r * 8;
        double* t1 = tile_ptr(band, I0 + k0 + 4, Jp) + r * 8;
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            double2 o;
            o.x = (lane < 8 && c >= lane) ? (c == lane ? rdiag[c] : a0[c]) : 0.0;
            o.y = (lane < 8 && c + 1 >= lane) ? (c + 1 == lane ? rdiag[c + 1] : a0[c + 1]) : 0.0;
            *reinterpret_cast<double2*>(t0 + c) = o;
            *reinterpret_cast<double2*>(t1 + c) = make_double2(0.0, 0.0);
        }
    }
    QR_TICK(4)
}

// R <- R * T for an accumulator-layout tile R (8 x 8) and the upper triangular T in shared memory
__device__ __forceinline__ void times_T(double& r0, double& r1, const double* Tn, int g, int q) {
    double c0 = 0.0, c1 = 0.0;
    dmma(c0, c1, r0, Tn[(2 * q) * VS + g]);
    dmma(c0, c1, r1, Tn[(2 * q + 1) * VS + g]);
    r0 = c0;
    r1 = c1;
}

// the next sweep's first panel can be factored during the last step of the sweep that starts at column c0: there is a next sweep with work and
// this sweep has at least two steps (the panel needs step 0's symmetric block and step 1's factored panel)
__device__ __forceinline__ bool sbr_lookahead(int N, int c0) { return c0 + 16 < N && (N - c0 - 8 + 63) / 64 >= 2; }

#ifdef FKMC_SBR_TIMING
#define SBR_TICK(i) { const long long now__ = clock64(); tk[i] += now__ - tq; tq = now__; }
#else
#define SBR_TICK(i)
#endif

__global__ void __launch_bounds__(SBR_THREADS, SBR_CTAS) sb2sb_kernel(double* band_all, size_t mat_stride, int N) {
    extern __shared__ __align__(16) unsigned char sbr_raw[];
    sbr_smem& sm = *reinterpret_cast<sbr_smem*>(sbr_raw);
    double* band = band_all + (size_t)blockIdx.x * mat_stride;
    const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler: the role branches below are convergent
    // roles from the hardware warp slots: 8 = panel, 0..7 = compute row block, -1 = idle
    __shared__ int s_slot[SBR_WARPS];
    {
        unsigned hw;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
        if (lane == 0) s_slot[wid] = (int)(hw & 3u);
    }
    __syncthreads();
    int w;
    {
        int n0 = 0, before0 = 0, before_other = 0;
        for (int i = 0; i < SBR_WARPS; ++i) {
            const bool z = s_slot[i] == 0;
            n0 += z;
            if (i < wid) { before0 += z; before_other += !z; }
        }
        const int n_other = SBR_WARPS - n0;
        if (n0 >= 1 && n_other + n0 - 1 >= 8) {
            // panel: the first warp of sub-partition 0; compute: the warps of the other sub-partitions first, then (if those are fewer
            // than eight) the remaining warps of sub-partition 0
            if (s_slot[wid] == 0) {
                const int k = n_other + before0 - 1;   // rank among the compute candidates for the second, third ... warp of sub-partition 0
                w = (before0 == 0) ? 8 : (k < 8 ? k : -1);
            } else {
                w = before_other < 8 ? before_other : -1;
            }
        } else {
            w = wid < 9 ? wid : -1;   // unexpected slot pattern: plain assignment (slower, still correct)
        }
        w = __shfl_sync(0xffffffffu, w, 0);
    }
    const int nsweeps = (N + 7) / 8;
    const int cswz = (2 * q) ^ (((g >> 1) & 1) << 2);       // accumulator-layout column pair inside a swizzled tile
#ifdef FKMC_SBR_TIMING
    long long tk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq = 0, nst = 0, qt[5] = {0, 0, 0, 0, 0};
#endif

    if (w < 0) {
        for (int j = 0; j < nsweeps; ++j) {
            if (8 * j + 8 >= N) break;
            sweep_barrier();
        }
        return;
    }
    if (w == 8) {
        // ================= panel warp =================
        int vpar = 0;
        bool ahead = false;   // the first panel of this sweep was factored at the end of the previous one
        for (int j = 0; j < nsweeps; ++j) {
            const int c0 = 8 * j;
            if (c0 + 8 >= N) break;
            sweep_barrier();
            const int k0 = lane >> 3, r = lane & 7;
            for (int p = 0;; ++p, ++vpar) {
                const int r0 = c0 + 8 + 64 * p;
                if (r0 >= N) break;
                const int I0 = r0 >> 3;
#ifdef FKMC_SBR_TIMING
                tq = clock64(); ++nst;
#endif
                if (p == 0 && ahead) {
                    bar_arrive(1, SBR_SYNC);
                    continue;
                }
                double a0[8], a1[8];
                if (p == 0) {
                    const double* t0 = tile_ptr(band, I0 + k0, I0 - 1) + r * 8;
                    const double* t1 = tile_ptr(band, I0 + k0 + 4, I0 - 1) + r * 8;
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        const double2 x = *reinterpret_cast<const double2*>(t0 + c), y = *reinterpret_cast<const double2*>(t1 + c);
                        a0[c] = x.x; a0[c + 1] = x.y; a1[c] = y.x; a1[c + 1] = y.y;
                    }
                } else {
                    bar_sync(2, SBR_SYNC);
                    const double* t0 = sm.P + k0 * 64;
                    const double* t1 = sm.P + (k0 + 4) * 64;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        a0[c] = t0[swz(r, c)];
                        a1[c] = t1[swz(r, c)];
                    }
                }
                SBR_TICK(0)
#ifdef FKMC_SBR_TIMING
                qr_panel(a0, a1, sm.V[vpar & 1], sm.pad, band, I0, p == 0 ? I0 - 1 : I0 - 8, lane, true, qt);
#else
                qr_panel(a0, a1, sm.V[vpar & 1], sm.pad, band, I0, p == 0 ? I0 - 1 : I0 - 8, lane, true);
#endif
                SBR_TICK(1)
            }
            // Look-ahead: the first panel of the next sweep (columns c0+8 .. c0+15, rows c0+16 .. c0+79) is final once step 0 of this sweep
            // has stored its symmetric block and step 1 its factored panel, so it is factored now, while the compute warps apply this
            // sweep's last step, instead of at the start of the next sweep with every compute warp waiting.  The compute warps arrive
            // on barrier 2 in their last step: they are past the step that still read the reflector buffer written here.
            ahead = sbr_lookahead(N, c0);
            if (ahead) {
                bar_sync(2, SBR_SYNC);
                __threadfence_block();   // (the factored panel of step 1 was stored by other lanes of this warp)
                __syncwarp();
                const int I0 = (c0 + 16) >> 3;
                double a0[8], a1[8];
                const double* t0 = tile_ptr(band, I0 + k0, I0 - 1) + r * 8;
                const double* t1 = tile_ptr(band, I0 + k0 + 4, I0 - 1) + r * 8;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    const double2 x = *reinterpret_cast<const double2*>(t0 + c), y = *reinterpret_cast<const double2*>(t1 + c);
                    a0[c] = x.x; a0[c + 1] = x.y; a1[c] = y.x; a1[c + 1] = y.y;
                }
#ifdef FKMC_SBR_TIMING
                qr_panel(a0, a1, sm.V[vpar & 1], sm.pad, band, I0, I0 - 1, lane, false, qt);
#else
                qr_panel(a0, a1, sm.V[vpar & 1], sm.pad, band, I0, I0 - 1, lane, false);
#endif
            }
        }
#ifdef FKMC_SBR_TIMING
        if (blockIdx.x == 0 && lane == 0) printf("sb2sb panel warp: steps %lld wait+load %lld qr %lld (cycles per step); per QR: products+head %lld reduce %lld scalar %lld update+T+stores %lld tail %lld\n", nst, tk[0] / nst, tk[1] / nst, qt[0] / nst, qt[1] / nst, qt[2] / nst, qt[3] / nst, qt[4] / nst);
#endif
        return;
    }

    // ================= compute warps =================
    int vpar = 0;
    for (int j = 0; j < nsweeps; ++j) {
        const int c0 = 8 * j;
        if (c0 + 8 >= N) break;
        sweep_barrier();   // the previous sweep's global stores are visible to every warp of the CTA
        for (int p = 0;; ++p, ++vpar) {
            const int r0 = c0 + 8 + 64 * p;
            if (r0 >= N) break;
            const int I0 = r0 >> 3;
            const bool has_next = r0 + 64 < N;
            const vset& vs = sm.V[vpar & 1];
#ifdef FKMC_SBR_TIMING
            tq = clock64(); ++nst;
#endif
            // ---- operand prefetch: row block w of S (upper tiles transposed from the stored lower ones) and of G
            double2 S[8], G[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                // lower tiles as stored; an upper tile is the transpose of the stored tile (c, w): loaded like any tile (one 512-byte
                // access) and transposed below with two shuffles instead of two strided loads that touch all 16 sectors each
                S[c] = *reinterpret_cast<const double2*>((c <= w ? tile_ptr(band, I0 + w, I0 + c) : tile_ptr(band, I0 + c, I0 + w)) + 2 * lane);
                G[c] = *reinterpret_cast<const double2*>(tile_ptr(band, I0 + 8 + w, I0 + c) + 2 * lane);
            }
            {
                // transpose of an accumulator-layout tile: element (g, 2q + e) of the transpose = element (2q + e, g) of the tile, held by
                // lane (2q + e, g >> 1) in component g & 1
                const int src0 = ((2 * q) << 2) | (g >> 1), src1 = ((2 * q + 1) << 2) | (g >> 1);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c > w) {   // warp-uniform
                        const double ax = __shfl_sync(0xffffffffu, S[c].x, src0), ay = __shfl_sync(0xffffffffu, S[c].y, src0);
                        const double bx = __shfl_sync(0xffffffffu, S[c].x, src1), by = __shfl_sync(0xffffffffu, S[c].y, src1);
                        S[c].x = (g & 1) ? ay : ax;
                        S[c].y = (g & 1) ? by : bx;
                    }
                }
            }
            SBR_TICK(0)
            bar_sync(1, SBR_SYNC);   // V, T of this step ready; the B block in shared memory complete
            SBR_TICK(1)

            // ---- phase 1a: G <- G - (G V T) V^T in registers; its first tile column is the next panel
            {
                double u0[4] = {0.0, 0.0, 0.0, 0.0}, u1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    dmma(u0[(2 * c) & 3], u1[(2 * c) & 3], G[c].x, vs.Vp[(8 * c + q) * VS + g]);
                    dmma(u0[(2 * c + 1) & 3], u1[(2 * c + 1) & 3], G[c].y, vs.Vp[(8 * c + 4 + q) * VS + g]);
                }
                double ua = (u0[0] + u0[1]) + (u0[2] + u0[3]), ub = (u1[0] + u1[1]) + (u1[2] + u1[3]);
                times_T(ua, ub, vs.Tn, g, q);
                ua = -ua;
                ub = -ub;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    dmma(G[c].x, G[c].y, ua, vs.Vq[(8 * c + g) * VS + q]);
                    dmma(G[c].x, G[c].y, ub, vs.Vq[(8 * c + g) * VS + 4 + q]);
                    if (c == 0 && has_next) {
                        *reinterpret_cast<double2*>(sm.P + w * 64 + g * 8 + cswz) = G[0];
                        bar_arrive(2, SBR_SYNC);
                    } else if (c == 0 && sbr_lookahead(N, c0)) {
                        bar_arrive(2, SBR_SYNC);   // last step of the sweep: releases the panel warp's look-ahead factorisation
                    }
                }
            }
            SBR_TICK(2)
            // ---- phase 1b
            if (p > 0 && w > 0) {
                // (T^T V^T B)^T for column tile w: contraction over the 64 rows, both operands from shared memory
                double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double* bt = sm.G + (k * 8 + w) * 64;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
                        dmma(a0[(2 * k + kk) & 3], a1[(2 * k + kk) & 3], bt[swz(4 * kk + q, g)], vs.Vn[(8 * k + 4 * kk + q) * VS + g]);
                }
                double wa = (a0[0] + a0[1]) + (a0[2] + a0[3]), wb = (a1[0] + a1[1]) + (a1[2] + a1[3]);
                times_T(wa, wb, vs.Tn, g, q);
                *reinterpret_cast<double2*>(&sm.W[(8 * w + g) * VS + 2 * q]) = make_double2(wa, wb);
            }
            double x0, x1;
            {
                // X = S V T (rows of this warp), partial of (V^T X)^T
                double xa[4] = {0.0, 0.0, 0.0, 0.0}, xb[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    dmma(xa[(2 * c) & 3], xb[(2 * c) & 3], S[c].x, vs.Vp[(8 * c + q) * VS + g]);
                    dmma(xa[(2 * c + 1) & 3], xb[(2 * c + 1) & 3], S[c].y, vs.Vp[(8 * c + 4 + q) * VS + g]);
                }
                x0 = (xa[0] + xa[1]) + (xa[2] + xa[3]);
                x1 = (xb[0] + xb[1]) + (xb[2] + xb[3]);
                times_T(x0, x1, vs.Tn, g, q);
                *reinterpret_cast<double2*>(&sm.XY[(8 * w + g) * VS + 2 * q]) = make_double2(x0, x1);
                __syncwarp();
                double z0 = 0.0, z1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) dmma(z0, z1, sm.XY[(8 * w + 4 * kk + q) * VS + g], vs.Vn[(8 * w + 4 * kk + q) * VS + g]);
                *reinterpret_cast<double2*>(&sm.Zp[w][g * 8 + 2 * q]) = make_double2(z0, z1);
            }
            SBR_TICK(3)
            bar_sync(3, 256);   // W, Zp complete
            SBR_TICK(4)

            // ---- phase 2
            const double av0 = vs.Vn[(8 * w + g) * VS + q], av1 = vs.Vn[(8 * w + g) * VS + 4 + q];
            if (p > 0) {
#pragma unroll
                for (int c = 1; c < 8; ++c) {
                    double2 b = *reinterpret_cast<const double2*>(sm.G + (w * 8 + c) * 64 + g * 8 + cswz);
                    dmma(b.x, b.y, -av0, sm.W[(8 * c + g) * VS + q]);
                    dmma(b.x, b.y, -av1, sm.W[(8 * c + g) * VS + 4 + q]);
                    *reinterpret_cast<double2*>(tile_ptr(band, I0 + w, I0 - 8 + c) + 2 * lane) = b;
                }
            }
            {
                // M = (V^T X)^T T (symmetric), Y = X - V M / 2
                double z0 = 0.0, z1 = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double2 z = *reinterpret_cast<const double2*>(&sm.Zp[k][g * 8 + 2 * q]);
                    z0 += z.x;
                    z1 += z.y;
                }
                times_T(z0, z1, vs.Tn, g, q);
                // B operand M[k = q + 4kk][n = g] = M[g][q + 4kk]: element (q + 4kk) of row g sits in lane (g, (q + 4kk) / 2), component (q & 1)
                const int src0 = (g << 2) | (q >> 1), src1 = (g << 2) | (2 + (q >> 1));
                const double m00 = __shfl_sync(0xffffffffu, z0, src0), m01 = __shfl_sync(0xffffffffu, z1, src0);
                const double m10 = __shfl_sync(0xffffffffu, z0, src1), m11 = __shfl_sync(0xffffffffu, z1, src1);
                const double mb0 = (q & 1) ? m01 : m00, mb1 = (q & 1) ? m11 : m10;
                dmma(x0, x1, -0.5 * av0, mb0);
                dmma(x0, x1, -0.5 * av1, mb1);
                *reinterpret_cast<double2*>(&sm.XY[(8 * w + g) * VS + 2 * q]) = make_double2(x0, x1);
            }
            SBR_TICK(5)
            bar_sync(3, 256);   // Y complete; nobody reads the B block any more
            SBR_TICK(6)

            // ---- phase 3: the updated G becomes the next step's B block; S -= V Y^T + Y V^T (lower tiles of this row block), back to the band
            if (has_next) {
#pragma unroll
                for (int c = 1; c < 8; ++c) *reinterpret_cast<double2*>(sm.G + (w * 8 + c) * 64 + g * 8 + cswz) = G[c];
            }
            {
                const double ay0 = sm.XY[(8 * w + g) * VS + q], ay1 = sm.XY[(8 * w + g) * VS + 4 + q];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (c <= w) {
                        dmma(S[c].x, S[c].y, -av0, sm.XY[(8 * c + g) * VS + q]);
                        dmma(S[c].x, S[c].y, -av1, sm.XY[(8 * c + g) * VS + 4 + q]);
                        dmma(S[c].x, S[c].y, -ay0, vs.Vn[(8 * c + g) * VS + q]);
                        dmma(S[c].x, S[c].y, -ay1, vs.Vn[(8 * c + g) * VS + 4 + q]);
                        *reinterpret_cast<double2*>(tile_ptr(band, I0 + w, I0 + c) + 2 * lane) = S[c];
                    }
                }
            }
            SBR_TICK(7)
        }
    }
#ifdef FKMC_SBR_TIMING
    if (blockIdx.x == 0 && lane == 0 && (w == 0 || w == 1 || w == 7))
        printf("sb2sb compute warp %d steps %lld: prefetch %lld wait-V %lld ph1a %lld ph1b %lld wait %lld ph2 %lld wait %lld ph3 %lld (cycles per step)\n", w, nst,
               tk[0] / nst, tk[1] / nst, tk[2] / nst, tk[3] / nst, tk[4] / nst, tk[5] / nst, tk[6] / nst, tk[7] / nst);
#endif
}

// Block row I of every matrix: the static hopping part plus the diagonal U f - mu_c of the permuted sites.
__global__ void __launch_bounds__(256) band_build_kernel(const double* __restrict__ band0, const int* __restrict__ perm, const int32_t* __restrict__ f,
                                                         int N, double U, double mu_c, double* __restrict__ band_all, size_t mat_stride) {
    const int I = blockIdx.x, b = blockIdx.y;
    const double4* src = reinterpret_cast<const double4*>(band0 + (size_t)I * NT * 64);
    double4* dst = reinterpret_cast<double4*>(band_all + (size_t)b * mat_stride + (size_t)I * NT * 64);
    double4 v = src[threadIdx.x];
    // diagonal tile = the last of the 16; its diagonal elements are words r*9 of the tile: double4 index (15*64 + 9r)/4, component (9r) & 3
    const int word = threadIdx.x * 4 - (NT - 1) * 64;
    if (word >= 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int wd = word + k;
            if (wd % 9 == 0) {
                const int r = I * 8 + wd / 9;
                if (r < N) {
                    const double d = U * (double)f[(size_t)b * N + perm[r]] - mu_c;
                    (k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w) += d;
                }
            }
        }
    }
    dst[threadIdx.x] = v;
}

// Result (half-bandwidth 8) in the layout sb2st.cu reads: AB[b][dd][c] = A[c + dd][c]
__global__ void __launch_bounds__(256) band_to_ab_kernel(const double* __restrict__ band_all, size_t mat_stride, int N, double* __restrict__ AB_all) {
    const int b = blockIdx.y;
    const double* band = band_all + (size_t)b * mat_stride;
    double* AB = AB_all + (size_t)b * 9 * N;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 9 * N; idx += gridDim.x * blockDim.x) {
        const int dd = idx / N, c = idx % N, r = c + dd;
        double v = 0.0;
        if (r < N) {
            const int I = r >> 3, J = c >> 3;
            v = band[((size_t)I * NT + (J - I + NT - 1)) * 64 + (r & 7) * 8 + (c & 7)];
        }
        AB[idx] = v;
    }
}

}  // namespace

bool fkmc_use_band(const fkmc_ctx* ctx) { return ctx->band_path && ctx->band_bw > 0 && ctx->N >= ctx->band_min && ctx->N <= 1024; }

size_t fkmc_band_stride(int N) { return (size_t)((N + 7) / 8 + 16) * NT * 64; }

// Chooses the site ordering (identity or slow coordinate folded) with the smaller bandwidth; the band path is usable when that is <= 64.
int fkmc_band_setup(fkmc_ctx* ctx) {
    const int N = ctx->N, L = ctx->L, Z = ctx->Z;
    ctx->band_bw = 0;
    if (N < 8) return FKMC_OK;
    const int inner = N / L;   // the first coordinate is the slowest
    std::vector<int> best_pos;
    int best_bw = N;
    for (int fold = 0; fold < 3; ++fold) {
        // 0: identity; 1: slow coordinate folded (0, L-1, 1, L-2, ...); 2: folded, and the fast coordinate mirrored on the returning half
        // (keeps the diagonal bonds of the triangular lattice at distance 2L + 1 on both halves)
        std::vector<int> pos(N);   // pos[site] = position in the new ordering
        for (int i = 0; i < N; ++i) {
            const int y = i / inner;
            int rest = i % inner;
            const bool back = y >= (L + 1) / 2;
            const int phi = !fold ? y : (!back ? 2 * y : 2 * (L - 1 - y) + 1);
            if (fold == 2 && back) rest = rest - rest % L + (L - 1 - rest % L);
            pos[i] = phi * inner + rest;
        }
        int bw = 0;
        for (int z = 0; z < Z; ++z)
            for (int i = 0; i < N; ++i) {
                const int jn = ctx->h_nbr_idx[(size_t)z * N + i];
                if (jn < N && ctx->h_nbr_val[(size_t)z * N + i] != 0.0) bw = std::max(bw, std::abs(pos[i] - pos[jn]));
            }
        if (bw < best_bw) { best_bw = bw; best_pos = pos; }
    }
    if (best_bw > 64 || best_bw < 1) return FKMC_OK;
    const size_t stride = fkmc_band_stride(N);
    std::vector<double> band0(stride, 0.0);
    std::vector<int> perm(N);
    for (int i = 0; i < N; ++i) perm[best_pos[i]] = i;
    for (int z = 0; z < Z; ++z)
        for (int i = 0; i < N; ++i) {
            const int jn = ctx->h_nbr_idx[(size_t)z * N + i];
            const double v = ctx->h_nbr_val[(size_t)z * N + i];
            if (jn >= N || v == 0.0) continue;
            const int r = best_pos[i], c = best_pos[jn];
            // entry (r, c) = v (the dense build assigns, duplicates coincide); stored when it falls into a lower or diagonal tile
            const int I = r >> 3, J = c >> 3;
            if (I >= J) band0[((size_t)I * NT + (J - I + NT - 1)) * 64 + (r & 7) * 8 + (c & 7)] = v;
        }
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_band0, sizeof(double) * stride));
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_band_perm, sizeof(int) * N));
    FKMC_CUDA(ctx, cudaMemcpy(ctx->d_band0, band0.data(), sizeof(double) * stride, cudaMemcpyHostToDevice));
    FKMC_CUDA(ctx, cudaMemcpy(ctx->d_band_perm, perm.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
    ctx->band_bw = best_bw;
    return FKMC_OK;
}

// f (device) -> half-bandwidth-8 band in d_AB (the input of fkmc_launch_sb2st)
int fkmc_launch_band_reduce(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double* d_band, double* d_AB) {
    const int N = ctx->N;
    const size_t stride = fkmc_band_stride(N);
    const int NB = (N + 7) / 8 + 16;
    {
        fkmc_prof_scope ps(ctx, "band_build");
        band_build_kernel<<<dim3(NB, B), 256, 0, ctx->stream>>>(ctx->d_band0, ctx->d_band_perm, d_f, N, U, mu_c, d_band, stride);
        ctx->launches++;
    }
    {
        fkmc_prof_scope ps(ctx, "sb2sb");
        FKMC_CUDA(ctx, cudaFuncSetAttribute(sb2sb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(sbr_smem)));
        sb2sb_kernel<<<B, SBR_THREADS, sizeof(sbr_smem), ctx->stream>>>(d_band, stride, N);
        ctx->launches++;
    }
    {
        fkmc_prof_scope ps(ctx, "band_build");
        band_to_ab_kernel<<<dim3(8, B), 256, 0, ctx->stream>>>(d_band, stride, N, d_AB);
        ctx->launches++;
    }
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}
