// Stage 2 of the two-stage tridiagonalisation: symmetric band (half-bandwidth 8) -> tridiagonal by
// Householder bulge chasing, entirely in shared memory, one CTA per matrix.
//
// Together with sy2sb.cu this replaces the tridiagonalisation stage of Eigen::SelfAdjointEigenSolver
// as called from configuration_t::calc_ed (src/configuration.cpp:212-213).
//
// Storage: Wb[c][d] = A(c+d, c), d = 0..15 (band 0..8 plus room for the transient fill 9..15), i.e. each
// matrix column from its diagonal downward in 16 consecutive doubles.  Sweep j annihilates column j below
// the sub-diagonal with an 8-row reflector and chases the resulting bulge down the band in blocks of 8
// (Murata-Horikoshi / Lang).  Sweep j+1 may execute its step t once sweep j has finished step t+1: with p = j + 1 + 8t, step u of
// sweep j touches rows / columns [p + 8(u-t) - 7, p + 8(u-t) + 16) -- block (i) rows [P, P+8) x columns [P-7, P-1], block (ii) [P, P+8)^2,
// block (iii) rows [P+8, P+16) x columns [P, P+8) with P = p + 8(u-t) -- and step t of sweep j+1 the same sets shifted by one, i.e.
// columns < p + 9 and rows < p + 17.  For u = t+2 the three blocks of sweep j have columns >= p + 9 (block (i)) or >= p + 16, so they are
// disjoint from everything step t of sweep j+1 reads or writes; u = t+1 does overlap (its block (i) is rows [p+8, p+16) x columns [p+1, p+7]).
// Hence a lag of two steps (round 1 used three); tools/sb2st_lag_check.py compares the two builds bit for bit.
//
// Work decomposition: EIGHT LANES PER SWEEP, four consecutive sweeps per warp.  Every lane keeps the whole
// reflector (8 doubles) in registers and owns one column (left update of the previous block, two-sided update of
// the diagonal block) or one row (right update of the block below) of the 8x8 blocks, so all contractions with v
// run inside a lane; per step a group needs only two 8-value broadcasts and two 8-lane sums.  The four groups of
// a warp run in lockstep three steps apart -- exactly the lag the data dependence asks for -- so only the first
// group of a warp has to watch another warp's progress counter.
#include <cfloat>

#include "common.cuh"

namespace {

#ifndef FKMC_SB2ST_SLEEP
#define FKMC_SB2ST_SLEEP 20
#endif
constexpr int SB = 8;     // half bandwidth
#ifndef FKMC_SB2ST_WD
#define FKMC_SB2ST_WD 19
#endif
#ifndef FKMC_SB2ST_WD_BIG
#define FKMC_SB2ST_WD_BIG 24
#endif
constexpr int WD_SMALL = FKMC_SB2ST_WD;    // doubles per stored column: 16 sub-diagonals (band 0..8 + transient fill 9..15) padded to a stride that spreads the three
                          // access patterns of a step (and the four sweeps of a warp, 24 columns apart) over the banks: 78 wavefronts per step instead of 120
#ifndef FKMC_SB2ST_LAG
#define FKMC_SB2ST_LAG 2
#endif
// Stride of the one-CTA-per-SM instantiation.  A 64-bit shared access is served per half-warp, i.e. per pair of sweep groups, whose columns are
// 8 LAG - 1 = 15 apart: the eight consecutive doubles each group reads from the block below fall on the same banks unless 15 WD = 8 (mod 16),
// i.e. WD = 8 (mod 16).  Modelled wavefronts of the band accesses per warp-step: 142 at WD = 19 (ncu: 26 % excess), 108 at WD = 24, 96 ideal;
// measured 13.78 -> 12.33 ms per 1024 matrices at N = 1024, bit-identical spectra (tools/sb2st_wd_check.py).  206 KB at N = 1024, so only
// matrices that own an SM anyway take it.
constexpr int WD_BIG = FKMC_SB2ST_WD_BIG;
constexpr int LAG = FKMC_SB2ST_LAG;    // steps between consecutive sweeps (see the dependence analysis above; 3 = the conservative value of round 1)

// Ordering of the band updates against the progress counters.  Writer: band stores (all lanes), __syncwarp, counter store (one
// lane); reader: counter load, __syncwarp, band loads.  Both sides are plain shared-memory accesses of one SM, which the LSU
// performs in issue order, so a compiler barrier is all that is needed -- a __threadfence_block() here is a MEMBAR.SC.CTA twice per
// step, which measured at about half of the step's latency (the step chain is the critical path of the whole kernel).
__device__ __forceinline__ void sched_fence() { asm volatile("" ::: "memory"); }
// progress counters with release / acquire semantics at CTA scope (PTX memory model: the band stores of a step happen-before the
// band loads of the sweep that waits on its counter).  FKMC_SB2ST_RELAXED restores the plain volatile accesses.
__device__ __forceinline__ int prog_load(const int* p) {
    int v;
#ifdef FKMC_SB2ST_RELAXED
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
#else
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
#endif
    return v;
}
__device__ __forceinline__ void prog_store(int* p, int v) {
#ifdef FKMC_SB2ST_RELAXED
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#else
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#endif
}

// sum over the 8 lanes of a group (m: the group's lane mask)
// sum over the eight lanes of a sweep group.  Full-mask shuffles: the whole warp executes every step together (inactive groups run
// the same instructions on dummy data), which spares the MATCH / REDUX / VOTE sequence the compiler puts in front of every collective
// with a partial mask.
__device__ __forceinline__ double gsum8(double x) {
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    return x;
}

// Branch-free reciprocal square root and reciprocal (MUFU seed + Newton steps).  The library versions carry a slow-path branch,
// which ends the basic block: without it the scheduler can run the block updates of a step under the latency of this chain.
__device__ __forceinline__ double rsqrt_nb(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // the seed is good to 2^-20 (tools/ubench/seedacc.cu): one third-order step reaches 2.7e-16, two quadratic steps of the reciprocal are exact
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}
__device__ __forceinline__ double rcp_nb(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; ++it) y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

struct refl {
    double v;    // lane l of the group holds v_l; v_0 = 1
    double tau;
    double beta;
};

// Householder reflector for x (x_l in lane l of the group, l < n; zero beyond n)
__device__ __forceinline__ refl make_reflector(double x, int l, int n, int gbase) {
    refl R;
    const double tail2 = gsum8((l >= 1 && l < n) ? x * x : 0.0);
    const double x0 = __shfl_sync(0xffffffffu, x, gbase);
    const bool triv = tail2 <= DBL_MIN;
    // |beta| = sqrt(x0^2 + tail2) through one reciprocal square root; tau = (beta - x0) / beta = 1 + |x0| / |beta| needs no division
    const double s2 = fma(x0, x0, triv ? 1.0 : tail2);
    const double rs = rsqrt_nb(s2);
    const double nrm = s2 * rs;
    const double beta = (x0 >= 0.0) ? -nrm : nrm;
    const double tau = fma(fabs(x0), rs, 1.0);
    const double inv = rcp_nb(x0 - beta);
    R.beta = triv ? x0 : beta;
    R.tau = triv ? 0.0 : tau;
    R.v = (l == 0) ? 1.0 : ((l < n && !triv) ? x * inv : 0.0);
    return R;
}

// The same reflector with the group's x vector exchanged through its shared-memory pad (one store, four 128-bit loads, the norm
// summed in the lane) instead of three shuffle stages + one broadcast shuffle.  pad: 8 doubles, 16-byte aligned.
__device__ __forceinline__ refl make_reflector_pad(double x, int l, int n, double* pad) {
    refl R;
    pad[l] = x;
    __syncwarp();
    double xs[SB];
#pragma unroll
    for (int i = 0; i < SB; i += 2) {
        const double2 t = *reinterpret_cast<const double2*>(pad + i);
        xs[i] = t.x;
        xs[i + 1] = t.y;
    }
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int i = 1; i < SB; ++i) {
        const double xi = (i < n) ? xs[i] : 0.0;
        if (i & 1) t1 = fma(xi, xi, t1); else t0 = fma(xi, xi, t0);
    }
    const double tail2 = t0 + t1, x0 = xs[0];
    const bool triv = tail2 <= DBL_MIN;
    const double s2 = fma(x0, x0, triv ? 1.0 : tail2);
    const double rs = rsqrt_nb(s2);
    const double nrm = s2 * rs;
    const double beta = (x0 >= 0.0) ? -nrm : nrm;
    const double tau = fma(fabs(x0), rs, 1.0);
    const double inv = rcp_nb(x0 - beta);
    R.beta = triv ? x0 : beta;
    R.tau = triv ? 0.0 : tau;
    R.v = (l == 0) ? 1.0 : ((l < n && !triv) ? x * inv : 0.0);
    return R;
}

// BIG: one CTA per SM anyway (N > 700 or so: the band alone is more than half of the shared memory), up to twelve warps with as many registers
// as they like; otherwise at most eight warps and 128 registers, so that two or more CTAs share an SM.
template <bool BIG, int WD>
__global__ void __launch_bounds__(BIG ? 384 : 256, BIG ? 1 : 2)
sb2st_kernel(const double* __restrict__ AB_all, int N, double* __restrict__ d_all, double* __restrict__ e_all) {
    extern __shared__ double smem[];
    double* Wb = smem;                                         // [N + 16][16]
    int* prog = reinterpret_cast<int*>(Wb + (size_t)(N + 16) * WD);  // [N] steps completed per sweep
    // per-group broadcast pads (16 doubles: the reflector and the u vector of step (ii)): one store + four 128-bit loads per
    // lane instead of eight group-masked shuffles each
    double* bcast = Wb + ((((size_t)(N + 16) * WD + (N + 1) / 2 + 2)) & ~(size_t)1);  // (16-byte aligned)
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const double* AB = AB_all + (size_t)b * (SB + 1) * N;

    for (int idx = tid; idx < (N + 16) * WD; idx += T) {
        const int c = idx / WD, dd = idx % WD;
        Wb[idx] = (c < N && dd <= SB && c + dd < N) ? AB[(size_t)dd * N + c] : 0.0;
    }
    for (int j = tid; j < N; j += T) prog[j] = 0;
    __syncthreads();

    const int l = lane & 7, q = lane >> 3, gbase = lane & 24;
    double* bc = bcast + (size_t)(warp * 4 + q) * 24;  // [0..7] v, [8..15] u, [16..23] x
    const int nsweeps = N - 2;
    int doff[SB];   // element (i, l) of the symmetric diagonal block in column storage
#pragma unroll
    for (int i = 0; i < SB; ++i) doff[i] = min(i, l) * (WD - 1) + max(i, l);
    for (int jb = 4 * warp; jb < nsweeps; jb += 4 * nwarps) {
        const int j = jb + q;
        int p = j + 1;                  // first row of the current reflector's index set
        int n = min(SB, N - p);         // its size
        int s = -LAG * q;               // step of this group (negative: not started)
        bool done = j >= nsweeps;
        double v[SB], vl = 0.0, tau = 0.0;
#pragma unroll
        for (int i = 0; i < SB; ++i) v[i] = 0.0;
        while (!__all_sync(0xffffffffu, done)) {
            const bool act = !done && s >= 0;
            // only the first sweep of the warp depends on another warp: sweep jb-1 must have finished step s+2
            if (q == 0 && act && jb > 0) {
                while (prog_load(prog + jb - 1) < s + LAG) { __nanosleep(FKMC_SB2ST_SLEEP); }
            }
            __syncwarp();
            sched_fence();
            // Every lane runs the step; `act` only gates what a group stores and how its state advances (tau = 0, v = 0 in a group that
            // has not started, so its arithmetic is the identity on dummy loads).
            {
                const bool first = act && s == 0;
                if (__any_sync(0xffffffffu, first)) {
                    // ---- step 0: annihilate column j below the sub-diagonal ----
                    const double x = first ? Wb[j * WD + 1 + l] : 0.0;  // rows p..p+7 of column j (zero beyond the matrix)
                    const refl R = make_reflector(x, l, n, gbase);
                    if (first) {
                        if (l < n) Wb[j * WD + 1 + l] = (l == 0) ? R.beta : 0.0;
                        vl = R.v;
                        tau = R.tau;
                    }
                    bc[l] = vl;
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < SB; i += 2) {
                        const double2 t = *reinterpret_cast<const double2*>(bc + i);
                        v[i] = t.x;
                        v[i + 1] = t.y;
                    }
                    __syncwarp();
                }
                // One step = three independent block updates with the same reflector (tau = 0 makes all of them the identity):
                //   (i)   [s >= 1] left-apply H to columns p-7 .. p-1 of the bulge block A(J, .): lane l owns column p-8+l
                //   (ii)  two-sided update of the diagonal block D = A(J, J): lane l owns its (symmetric) column l
                //   (iii) right-apply H to the block below, A(Jn, J), Jn = rows p+8 .. p+15: lane l owns row l
                // All loads first, then the arithmetic in one basic block (the three FMA chains and the next reflector's
                // sqrt / divisions interleave), then the stores.
                const int pn = p + SB;
                const int nn = min(SB, N - pn);
                const bool c1 = act && (s >= 1) && (l >= 1);
                const bool c3 = act && (n == SB) && (nn > 0);
                double* col = Wb + (c1 ? (p - SB + l) : p) * WD + (SB - l);
                double* D0 = Wb + p * WD;
                double* B0 = Wb + p * WD + SB + l;
                double a[SB], dcol[SB], bel[SB];
#pragma unroll
                for (int i = 0; i < SB; ++i) {
                    bel[i] = c3 ? B0[i * (WD - 1)] : 0.0;
                    dcol[i] = act ? D0[doff[i]] : 0.0;
                    a[i] = c1 ? col[i] : 0.0;
                }
                // (iii) first: the next reflector hangs on it
                double z0 = 0.0, z1 = 0.0;
#pragma unroll
                for (int c = 0; c < SB; c += 2) {
                    z0 = fma(bel[c], v[c], z0);
                    z1 = fma(bel[c + 1], v[c + 1], z1);
                }
                const double z = tau * (z0 + z1);
#pragma unroll
                for (int c = 0; c < SB; ++c) bel[c] = fma(-z, v[c], bel[c]);
                const bool more = act && (n == SB && nn >= 2);
                // next reflector from the first column of the block below (rows pn.., column p): only this sweep touches it now
                const refl R = make_reflector_pad(bel[0], l, nn, bc + 16);
                // (i)
                double w0 = 0.0, w1 = 0.0;
#pragma unroll
                for (int i = 0; i < SB; i += 2) {
                    w0 = fma(v[i], a[i], w0);
                    w1 = fma(v[i + 1], a[i + 1], w1);
                }
                const double w = tau * (w0 + w1);
#pragma unroll
                for (int i = 0; i < SB; ++i) a[i] = fma(-v[i], w, a[i]);
                // (ii)
                double u0 = 0.0, u1 = 0.0;
#pragma unroll
                for (int i = 0; i < SB; i += 2) {
                    u0 = fma(dcol[i], v[i], u0);
                    u1 = fma(dcol[i + 1], v[i + 1], u1);
                }
                double u = tau * (u0 + u1);                        // u_l = tau (D v)_l
                // every lane reads the whole (uncorrected) u from the group's pad and forms alpha = -tau v^T u / 2 itself: one exchange
                // instead of a three-stage shuffle sum followed by the exchange of the corrected vector
                bc[8 + l] = u;
                __syncwarp();
                double ur[SB], q0 = 0.0, q1 = 0.0;
#pragma unroll
                for (int i = 0; i < SB; i += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(bc + 8 + i);
                    ur[i] = t.x;
                    ur[i + 1] = t.y;
                    q0 = fma(t.x, v[i], q0);
                    q1 = fma(t.y, v[i + 1], q1);
                }
                const double alpha = -0.5 * tau * (q0 + q1);
                u = fma(alpha, vl, u);
#pragma unroll
                for (int i = 0; i < SB; ++i) dcol[i] = dcol[i] - v[i] * u - fma(alpha, v[i], ur[i]) * vl;
                // stores
#pragma unroll
                for (int i = 0; i < SB; ++i) {
                    if (c1) col[i] = a[i];
                    if (act && i >= l) D0[l * (WD - 1) + i] = dcol[i];
                    if (c3) B0[i * (WD - 1)] = bel[i];
                }
                if (more && l < nn) Wb[p * WD + SB + l] = (l == 0) ? R.beta : 0.0;
                vl = more ? R.v : vl;
                tau = more ? R.tau : tau;
                bc[l] = vl;     // (a group that does not advance broadcasts the reflector it already holds)
                __syncwarp();
#pragma unroll
                for (int i = 0; i < SB; i += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(bc + i);
                    v[i] = t.x;
                    v[i + 1] = t.y;
                }
                p = more ? pn : p;
                n = more ? nn : n;
                done = done || (act && !more);
            }
            // publish progress: the writes of this step are visible before the counter moves
            __syncwarp();
            sched_fence();
            if (act && l == 0) prog_store(prog + j, done ? (1 << 30) : s + 1);
            ++s;
        }
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

// band, counters, broadcast pads (24 doubles per sweep group, four groups per warp)
static size_t sb2st_smem(int N, int nwarps, int WD = WD_SMALL) { return sizeof(double) * ((size_t)(N + 16) * WD + (N + 1) / 2 + 3 + (size_t)nwarps * 4 * 24) + 16; }
size_t fkmc_sb2st_smem(int N) { return sb2st_smem(N, 12); }  // upper bound (at most twelve warps)

int fkmc_launch_sb2st(fkmc_ctx* ctx, const double* d_AB, int N, int B, double* d_d, double* d_e) {
    fkmc_prof_scope ps(ctx, "sb2st");
    // sweeps in flight <= blocks along the band / lag, four sweeps per warp (measured with tools/sb2st_warps_scan.py: N = 1024 flat from 8
    // to 16 warps, N = 576 best at 8 with two CTAs per SM, N = 256 best at 3)
    const int nblk = (N + SB - 1) / SB;
    const bool big = sb2st_smem(N, 12) > 115712;   // more than half of an SM's shared memory: one CTA per SM whatever the warp count
    const int cap = big ? 12 : 8;
    int nwarps = ((nblk + LAG - 1) / LAG + 3) / 4 - 1;
    if (nwarps < 1) nwarps = 1;
    if (nwarps > cap) nwarps = cap;
    if (ctx->sb2st_warps > 0) nwarps = std::min(ctx->sb2st_warps, cap);  // tuning override (fkmc_set_option "sb2st_warps")
    // matrices that own an SM anyway take the conflict-free column stride WD_BIG when the wider band still fits (N <= 1122), else the compact one
    const bool wide = big && WD_BIG != WD_SMALL && sb2st_smem(N, nwarps, WD_BIG) <= ctx->smem_optin;
    const size_t smem = sb2st_smem(N, nwarps, wide ? WD_BIG : WD_SMALL);  // small matrices share an SM: no more shared memory than this launch needs
    if (smem > ctx->smem_optin) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sb2st: matrix too large for shared memory");
    if (wide) {
        FKMC_CUDA(ctx, cudaFuncSetAttribute(sb2st_kernel<true, WD_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb2st_kernel<true, WD_BIG><<<B, nwarps * 32, smem, ctx->stream>>>(d_AB, N, d_d, d_e);
    } else if (big) {
        FKMC_CUDA(ctx, cudaFuncSetAttribute(sb2st_kernel<true, WD_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb2st_kernel<true, WD_SMALL><<<B, nwarps * 32, smem, ctx->stream>>>(d_AB, N, d_d, d_e);
    } else {
        FKMC_CUDA(ctx, cudaFuncSetAttribute(sb2st_kernel<false, WD_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sb2st_kernel<false, WD_SMALL><<<B, nwarps * 32, smem, ctx->stream>>>(d_AB, N, d_d, d_e);
    }
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}
