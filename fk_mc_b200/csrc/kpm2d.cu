// Two-kernel Chebyshev / KPM weight evaluation for the regular 2-D lattices (cubic2d, triangular, honeycomb brick wall).
//
// Same contract as kpm.cu (configuration_t::calc_chebyshev, src/configuration.cpp:94-205; chebyshev_eval,
// include/fk_mc/chebyshev.hpp:21-54), split by what bounds each half:
//
//   lanczos2d_kernel   e_min / e_max of H (reference: two ARPACK solves, configuration.cpp:99-100).  Latency-bound, so the
//                      CTA is small (128 threads, one 1 x 8 strip of the lattice per thread) and many proposals share an SM.
//                      The stencil reads whole neighbour strips with 128-bit shared-memory loads, x neighbours stay in
//                      registers.  Ritz values of the Lanczos tridiagonal come from Laguerre's iteration on the
//                      characteristic polynomial (monotone from outside the spectrum, cubic convergence), confirmed and
//                      rounded by one 32-point Sturm count around the result; the multisection of kpm.cu is the fallback.
//   kpm_moments2d_kernel  exact full-trace moments mu_m = Tr T_m(X)/N by the column recursion (configuration.cpp:117-194).
//                      T_m(X) e_j lives on the sites within m hops of j, so the sites of the H-hop neighbourhood are numbered
//                      ring by ring (element e of a patch = lane e % 32, slot e / 32) and step m only touches the
//                      slots below cnt(m): 18 warp-wide element updates per column for H = 8 instead of 70 for the
//                      (2H+1)^2 square.  Neighbour positions come from a per-lattice table built on the host from the
//                      context's own stencil, so one kernel serves all three lattices.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <queue>

#include "common.cuh"

namespace {

constexpr int LZ_KMAX = 384;      // Lanczos step cap (as kpm.cu)
constexpr int LZ_T = 128;         // threads per proposal
// first Ritz evaluation at step max(64, 3 L): the number of steps both ends need grows like L (N = 1024: 112 .. 192, mean 135), and an
// evaluation that cannot yet confirm convergence only costs time (first evaluation at 64 for every size: 0.449 ms per 1024 proposals at
// L = 32, at 96: 0.418 ms)
constexpr int LZ_EVERY = 16;      // then every LZ_EVERY steps, compared with the previous evaluation

struct lz_args {
    const int32_t* f;
    int N, L;
    double U, mu_c, ht, hp;
    double* ab;   // [B][4]: e_min, e_max written here
    int* flag;
    int* steps;
    const int* order;  // [B] proposal handled by CTA i: longest expected run first (null: identity)
    int kmax;          // Lanczos step cap (LZ_KMAX unless the "lanczos_max_steps" option lowers it)
};

// ---------------------------------------------------------------------------------------------------------------------
// Sturm count of the k x k Lanczos tridiagonal: number of eigenvalues below x.  ab[i] = (alpha_i, beta_i^2), beta_0 = 0.
__device__ __forceinline__ int sturm_below(const double2* __restrict__ ab, int k, double x) {
    // p_{i+1} = (alpha_i - x) p_i - beta_i^2 p_{i-1}; an exact zero counts as negative (the next value then has the sign
    // opposite to the previous one, so the number of sign changes comes out right) -- the sign logic stays off the FMA chain
    double pm1 = 1.0, p = ab[0].x - x;
    bool neg = !(p > 0.0);
    int cnt = neg ? 1 : 0;
    int i = 1;
    for (; i + 8 <= k; i += 8) {
        double2 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = ab[i + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double pn = fma(q[u].x - x, p, -(q[u].y * pm1));
            const bool nneg = !(pn > 0.0);
            cnt += (nneg != neg) ? 1 : 0;
            neg = nneg;
            pm1 = p;
            p = pn;
        }
        const double m = fmax(fabs(p), fabs(pm1));
        if (m > 1.157920892373162e77) { p *= 8.636168555094445e-78; pm1 *= 8.636168555094445e-78; }
        else if (m < 8.636168555094445e-78) { p *= 1.157920892373162e77; pm1 *= 1.157920892373162e77; }
    }
    for (; i < k; ++i) {
        const double2 q = ab[i];
        const double pn = fma(q.x - x, p, -(q.y * pm1));
        const bool nneg = !(pn > 0.0);
        cnt += (nneg != neg) ? 1 : 0;
        neg = nneg;
        pm1 = p;
        p = pn;
    }
    return cnt;
}

// idx-th eigenvalue by 32-way multisection inside [lo, hi] (whole warp; the rigorous fallback)
__device__ __forceinline__ double warp_multisect(const double2* __restrict__ ab, int k, int idx, double lo, double hi, int lane) {
    const double pad = 8.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) + DBL_MIN;
    double a = lo - pad, c = hi + pad;
    for (int round = 0; round < 14; ++round) {
        const double h = (c - a) * (1.0 / 33.0);
        if (!(h > 2.0 * DBL_EPSILON * fmax(fabs(a), fabs(c)) * (1.0 / 33.0))) break;
        const double x = a + h * (double)(lane + 1);
        const bool above = sturm_below(ab, k, x) > idx;
        const unsigned mask = __ballot_sync(0xffffffffu, above);
        const int first = mask ? (__ffs(mask) - 1) : 32;
        const double na = first > 0 ? a + h * (double)first : a;
        const double nc = first < 32 ? a + h * (double)(first + 1) : c;
        a = na;
        c = nc;
    }
    return 0.5 * (a + c);
}

// p(x) = det(T_k - x) with p', p'' by a warp-wide product of 2x2 transfer matrices: (p_i, p_{i-1})^T = M_i (p_{i-1}, p_{i-2})^T,
// M_i = [[alpha_{i-1} - x, -beta_{i-1}^2], [1, 0]].  Lane l multiplies its chunk of rows, a butterfly over the lanes multiplies the
// chunks in order (the partner with the higher lane index holds the later rows = the left factor).  The triple (P, P', P'') is
// rescaled by a power of two after every stage: only the ratios p'/p and p''/p are used.
struct tm3 {
    double p[4], d[4], s[4];  // row-major 2x2: P, dP/dx, d2P/dx2
};
__device__ __forceinline__ void mm2(const double* A, const double* B, double* C) {  // C = A B
    C[0] = fma(A[0], B[0], A[1] * B[2]);
    C[1] = fma(A[0], B[1], A[1] * B[3]);
    C[2] = fma(A[2], B[0], A[3] * B[2]);
    C[3] = fma(A[2], B[1], A[3] * B[3]);
}
__device__ __forceinline__ void mm2acc(const double* A, const double* B, double* C, double w) {  // C += w A B
    C[0] = fma(w * A[0], B[0], fma(w * A[1], B[2], C[0]));
    C[1] = fma(w * A[0], B[1], fma(w * A[1], B[3], C[1]));
    C[2] = fma(w * A[2], B[0], fma(w * A[3], B[2], C[2]));
    C[3] = fma(w * A[2], B[1], fma(w * A[3], B[3], C[3]));
}
__device__ __forceinline__ void warp_charpoly3(const double2* __restrict__ ab, int k, double x, int lane, double& p, double& dp, double& ddp) {
    const int c = (k + 31) >> 5;
    tm3 t;
    t.p[0] = 1.0; t.p[1] = 0.0; t.p[2] = 0.0; t.p[3] = 1.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) t.d[i] = t.s[i] = 0.0;
    const int r0 = lane * c, r1 = min(k, r0 + c);
    for (int i = r0; i < r1; ++i) {
        const double2 q = ab[i];
        const double a = q.x - x, b = q.y;
        // X <- M X;  X' <- M X' - E11 X;  X'' <- M X'' - 2 E11 X'
        const double s0 = fma(a, t.s[0], fma(-b, t.s[2], -2.0 * t.d[0])), s1 = fma(a, t.s[1], fma(-b, t.s[3], -2.0 * t.d[1]));
        t.s[2] = t.s[0]; t.s[3] = t.s[1]; t.s[0] = s0; t.s[1] = s1;
        const double d0 = fma(a, t.d[0], fma(-b, t.d[2], -t.p[0])), d1 = fma(a, t.d[1], fma(-b, t.d[3], -t.p[1]));
        t.d[2] = t.d[0]; t.d[3] = t.d[1]; t.d[0] = d0; t.d[1] = d1;
        const double p0 = fma(a, t.p[0], -(b * t.p[2])), p1 = fma(a, t.p[1], -(b * t.p[3]));
        t.p[2] = t.p[0]; t.p[3] = t.p[1]; t.p[0] = p0; t.p[1] = p1;
    }
#pragma unroll
    for (int mask = 1; mask < 32; mask <<= 1) {
        tm3 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o.p[i] = __shfl_xor_sync(0xffffffffu, t.p[i], mask);
            o.d[i] = __shfl_xor_sync(0xffffffffu, t.d[i], mask);
            o.s[i] = __shfl_xor_sync(0xffffffffu, t.s[i], mask);
        }
        const bool high = (lane & mask) != 0;  // own chunk holds the later rows
        tm3 H, Lo;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            H.p[i] = high ? t.p[i] : o.p[i]; H.d[i] = high ? t.d[i] : o.d[i]; H.s[i] = high ? t.s[i] : o.s[i];
            Lo.p[i] = high ? o.p[i] : t.p[i]; Lo.d[i] = high ? o.d[i] : t.d[i]; Lo.s[i] = high ? o.s[i] : t.s[i];
        }
        mm2(H.p, Lo.p, t.p);
        mm2(H.d, Lo.p, t.d);
        mm2acc(H.p, Lo.d, t.d, 1.0);
        mm2(H.s, Lo.p, t.s);
        mm2acc(H.d, Lo.d, t.s, 2.0);
        mm2acc(H.p, Lo.s, t.s, 1.0);
        double m = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) m = fmax(m, fmax(fabs(t.p[i]), fmax(fabs(t.d[i]), fabs(t.s[i]))));
        const int ex = (__double2hiint(m) >> 20) & 0x7ff;
        const double sc = __hiloint2double(min(max(2046 - ex, 1), 2046) << 20, 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) { t.p[i] *= sc; t.d[i] *= sc; t.s[i] *= sc; }
    }
    p = t.p[0];
    dp = t.d[0];
    ddp = t.s[0];
}

// Laguerre's iteration on p(x) = det(T_k - x) from a start outside the spectrum (x0 below every eigenvalue for the lower
// end, above for the upper end): the iterates move monotonically towards the nearest root and never pass it.  Whole warp.
__device__ __forceinline__ double laguerre_end(const double2* __restrict__ ab, int k, double x, double tol, int lane, int& iters) {
    const double n = (double)k;
    for (int it = 0; it < 24; ++it) {
        ++iters;
        double p1, d1, s1;
        warp_charpoly3(ab, k, x, lane, p1, d1, s1);
        if (p1 == 0.0) return x;
        const double Gq = d1 / p1, Hq = Gq * Gq - s1 / p1;
        double disc = (n - 1.0) * (n * Hq - Gq * Gq);
        disc = disc > 0.0 ? sqrt(disc) : 0.0;
        const double den = Gq >= 0.0 ? Gq + disc : Gq - disc;
        if (den == 0.0) return x;
        const double a = n / den;
        x -= a;
        if (fabs(a) <= tol) break;
    }
    return x;
}

// Ritz value at one end of T_k (upper = false: smallest, true: largest), whole warp.  gl / gh enclose the spectrum.
// With a previous value (of a leading sub-matrix: interlacing puts it inside the new spectrum, next to the wanted end) the
// iteration starts there and typically needs two steps; a start outside the spectrum is the safe restart.  Either way the
// result is accepted only when 32 Sturm counts around it bracket the wanted eigenvalue; the multisection is the last resort.
__device__ __noinline__ double warp_ritz_end(const double2* __restrict__ ab, int k, bool upper, double gl, double gh, double hscale,
                                             int lane, bool have_prev, double prev, int& iters, int& fallbacks) {
    const double pad = 8.0 * DBL_EPSILON * fmax(fabs(gl), fabs(gh)) + DBL_MIN;
    const double delta = 2.0 * DBL_EPSILON * hscale;
    const int idx = upper ? k - 1 : 0;
    for (int attempt = have_prev ? 0 : 1; attempt < 2; ++attempt) {
        const double x0 = attempt == 0 ? prev : (upper ? gh + pad : gl - pad);
        const double xs = __shfl_sync(0xffffffffu, laguerre_end(ab, k, x0, 1e-9 * hscale, lane, iters), 0);
        // the eigenvalue sits between the last point with count <= idx and the first with count > idx
        const double x = xs + delta * ((double)lane - 15.5);
        const bool above = sturm_below(ab, k, x) > idx;
        const unsigned mask = __ballot_sync(0xffffffffu, above);
        if (mask != 0u && mask != 0xffffffffu) {
            const int first = __ffs(mask) - 1;  // counts are monotone in x: lanes >= first are above
            return xs + delta * ((double)first - 16.0);
        }
        ++fallbacks;
    }
    return warp_multisect(ab, k, idx, gl, gh, lane);
}

// ---------------------------------------------------------------------------------------------------------------------
// Lanczos for e_min / e_max.  Thread t owns the strip (y, 8 sx .. 8 sx + 7), t = y (L/8) + sx; site index = y L + x
// (hypercubic_lattice::pos_to_index: last coordinate fastest).
template <int KIND>
__global__ void __launch_bounds__(LZ_T, 4) lanczos2d_kernel(lz_args P) {
    extern __shared__ __align__(16) double sm[];
    const int N = P.N, L = P.L;
    const int lz_first = max(64, 3 * L);
    const int b = P.order ? P.order[blockIdx.x] : blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = LZ_T / 32;
    double* vb0 = sm;                                         // [N] v_k, ping
    double* vb1 = vb0 + N;                                    // [N] v_k, pong
    double2* ab = reinterpret_cast<double2*>(vb1 + N);        // [LZ_KMAX + 1] (alpha_i, beta_i^2)
    double* rpA = reinterpret_cast<double*>(ab + LZ_KMAX + 2);  // [NWARP]
    double* rpB = rpA + NWARP;                                // [NWARP]
    double* msc = rpB + NWARP;                                // [8]

    const int nstrip = L >> 3, NS = N >> 3;
    const bool active = tid < NS;
    const int y = active ? tid / nstrip : 0, sx = active ? tid - y * nstrip : 0, x0 = sx << 3;
    // shared vectors are strip-major: element e of strip t at e * NS + t, so that a warp's accesses are consecutive words
    const int tu = (y + 1 == L) ? sx : tid + nstrip, td = (y == 0) ? tid + NS - nstrip : tid - nstrip;     // strips above / below
    const int sl = (sx == 0) ? nstrip - 1 : -1, sr = (sx + 1 == nstrip) ? 1 - nstrip : 1;                      // strip offsets left / right
    const int base = y * L + x0;
    const bool yodd = (y & 1) != 0;

    double xd[8], vr[8], pr[8];
    {
        const int4* fp = reinterpret_cast<const int4*>(P.f + (size_t)b * N + base);
        int fv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (active) {
            const int4 a = fp[0], c = fp[1];
            fv[0] = a.x; fv[1] = a.y; fv[2] = a.z; fv[3] = a.w; fv[4] = c.x; fv[5] = c.y; fv[6] = c.z; fv[7] = c.w;
        }
        double part = 0.0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            xd[e] = P.U * (double)fv[e] - P.mu_c;
            double v = 0.0;
            if (active) {
                unsigned h = (unsigned)(base + e) * 2654435761u + 0x9e3779b9u;
                h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
                v = (double)h * (1.0 / 4294967296.0) - 0.5;
            }
            vr[e] = v;
            pr[e] = 0.0;
            part = fma(v, v, part);
        }
        part = warp_sum(part);
        if (lane == 0) rpA[warp] = part;
        if (tid == 0) ab[0].y = 0.0;
        __syncthreads();
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) s += rpA[w];
        const double inv = rsqrt(s);
#pragma unroll
        for (int e = 0; e < 8; ++e) vr[e] *= inv;
        if (active) {
#pragma unroll
            for (int e = 0; e < 8; ++e) vb0[e * NS + tid] = vr[e];
        }
        __syncthreads();
    }

    const int kcap = min(P.kmax, N);
    double e_min = 0.0, e_max = 0.0, gl = DBL_MAX, gh = -DBL_MAX, hscale = 0.0;
    double prev_min = 0.0, prev_max = 0.0;
    bool have_prev = false, converged = false;
    int k = 0;
    double beta_k = 0.0;
    const double ht = P.ht, hp = P.hp;
#ifdef FKMC_LZ_TIMING
    long long t_ritz = 0, t_start = clock64();
#endif
    int lag_iters = 0, lag_fallbacks = 0, nchecks = 0;
    while (k < kcap) {
        const double* lv = (k & 1) ? vb1 : vb0;
        double* lvn = (k & 1) ? vb0 : vb1;
        double pa = 0.0;
        if (active) {
            double u[9], d[9];  // u[e] = v(y+1, x0+e), e = 0..8;  d[e + 1] = v(y-1, x0+e), d[0] = v(y-1, x0-1)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                u[e] = lv[e * NS + tu];
                d[e + 1] = lv[e * NS + td];
            }
            const double left = lv[7 * NS + tid + sl], right = lv[tid + sr];
            if (KIND == FKMC_TRIANGULAR) {
                u[8] = lv[tu + sr];
                d[0] = lv[7 * NS + td + sl];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const double lf = (e == 0) ? left : vr[e - 1];
                const double rg = (e == 7) ? right : vr[e + 1];
                double nb = lf + rg;
                if (KIND == FKMC_HONEYCOMB) {
                    // sublattice A ((y + x) even) hops up, B hops down; x0 is a multiple of 8
                    const bool odd = ((e & 1) != 0) != yodd;
                    nb += odd ? d[e + 1] : u[e];
                } else {
                    nb += u[e] + d[e + 1];
                }
                double sacc = fma(xd[e], vr[e], ht * nb);
                if (KIND == FKMC_TRIANGULAR) sacc = fma(hp, d[e] + u[e + 1], sacc);
                sacc = fma(-beta_k, pr[e], sacc);
                pr[e] = sacc;  // w (v_{k-1} is not needed any more)
                pa = fma(sacc, vr[e], pa);
            }
        }
        pa = warp_sum(pa);
        if (lane == 0) rpA[warp] = pa;
        __syncthreads();
        double alpha = 0.0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) alpha += rpA[w];
        double pb = 0.0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            pr[e] = fma(-alpha, vr[e], pr[e]);
            pb = fma(pr[e], pr[e], pb);
        }
        pb = warp_sum(pb);
        if (lane == 0) rpB[warp] = pb;
        __syncthreads();
        double nb2 = 0.0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) nb2 += rpB[w];
        nb2 = fmax(nb2, 0.0);
        const double rinv = nb2 > 0.0 ? rsqrt(nb2) : 0.0;
        const double nbv = nb2 * rinv;
        if (tid == 0) { ab[k].x = alpha; ab[k + 1].y = nb2; }
        gl = fmin(gl, alpha - beta_k - nbv);
        gh = fmax(gh, alpha + beta_k + nbv);
        hscale = fmax(hscale, fabs(alpha) + nbv);
        ++k;
        const bool breakdown = nbv <= 1e-13 * hscale;
        if (!breakdown) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const double vn = pr[e] * rinv;
                pr[e] = vr[e];
                vr[e] = vn;
            }
            if (active) {
#pragma unroll
                for (int e = 0; e < 8; ++e) lvn[e * NS + tid] = vr[e];
            }
        }
        beta_k = nbv;
        __syncthreads();
        const bool last = breakdown || k == kcap;
        if (last || (k >= lz_first && ((k - lz_first) % LZ_EVERY) == 0)) {
            // extreme Ritz values of T_k (beta_k is outside T_k, which only loosens the Gershgorin enclosure); converged
            // when both ends have stopped moving since the previous evaluation
#ifdef FKMC_LZ_TIMING
            const long long t_r0 = clock64();
#endif
            ++nchecks;
            if (warp < 2) {
                const double v = warp_ritz_end(ab, k, warp == 1, gl, gh, hscale, lane, have_prev, warp == 1 ? prev_max : prev_min, lag_iters, lag_fallbacks);
                if (lane == 0) msc[warp] = v;
            }
            __syncthreads();
            e_min = msc[0];
            e_max = msc[1];
            const double tol = 8.0 * DBL_EPSILON * hscale;
            if (have_prev && fabs(e_min - prev_min) <= tol && fabs(e_max - prev_max) <= tol) converged = true;
            prev_min = e_min;
            prev_max = e_max;
            have_prev = true;
            __syncthreads();
#ifdef FKMC_LZ_TIMING
            t_ritz += clock64() - t_r0;
#endif
            if (converged || last) break;
        }
    }
#ifdef FKMC_LZ_TIMING
    if ((tid == 0 || tid == 32) && b < 4)
        printf("lanczos2d b=%d warp=%d: total %lld cycles, ritz %lld, steps %d, checks %d, laguerre iterations %d, fallbacks %d\n", b, warp,
               (long long)clock64() - t_start, t_ritz, k, nchecks, lag_iters, lag_fallbacks);
#endif
    if (tid == 0) {
        if (!converged && k == kcap && k < N) atomicOr(P.flag, 2);
        if (P.steps) P.steps[b] = k;
        P.ab[(size_t)b * 4 + 0] = e_min;
        P.ab[(size_t)b * 4 + 1] = e_max;
    }
}

// Launch order of the Lanczos CTAs: the number of steps a proposal needed in the previous launch predicts what it needs now (a chain's
// configuration changes by one site per proposal), and the runs differ by up to 2x, so the CTAs go out longest first (counting sort
// by the previous step count; one CTA).  The first launch sees zeros and keeps the identity order.
__global__ void __launch_bounds__(1024) lanczos_order_kernel(const int* __restrict__ steps, int B, int* __restrict__ order) {
    __shared__ int hist[64], start[64];
    const int tid = threadIdx.x;
    if (tid < 64) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < B; i += blockDim.x) atomicAdd(&hist[63 - min(63, max(0, steps[i]) >> 3)], 1);
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int k = 0; k < 64; ++k) { start[k] = acc; acc += hist[k]; }
    }
    __syncthreads();
    // the order inside one key is irrelevant (every proposal is computed independently of where its CTA runs)
    for (int i = tid; i < B; i += blockDim.x) order[atomicAdd(&start[63 - min(63, max(0, steps[i]) >> 3)], 1)] = i;
}

// ---------------------------------------------------------------------------------------------------------------------
// Moments.  The patch of column j (the sites within H hops of j) is kept in shared memory as a (2H+3) x P array addressed
// by relative position, a(dy, dx) = (dy + H + 1) P + (dx + H + 1), with a ring of cells around it that stay zero, so a neighbour
// is always "own address + constant" and nothing outside the patch needs a special case.  Every patch element is owned by one
// (slot, lane): elements are dealt out by increasing hop distance so that step m only has to visit the first ns(m) slots, and
// the lane inside a half-warp IS the address modulo 16 -- sixteen lanes of a half-warp therefore hit sixteen different bank
// pairs for their own cell and, shifted by the same constant, for every neighbour: no bank conflicts by construction.
// Tables (device, built by the host from the context's stencil), per parity class c, EP = 32 S entries (index 32 slot + lane):
//   tabi[c][m], m = 0..H    ns(m): slots that hold elements within m hops;   tabi[c][H+1] = lane of the centre,
//   tabi[c][H+2] = P,  tabi[c][H+3] = cells per buffer
//   off[c][e]               (hops << 16) | (dy + 128) << 8 | (dx + 128); hops = 255 for an unused (slot, lane)
//   nb[c][z][e], z < Z      cell of the neighbour through stencil slot z;   nb[c][Z][e] = own cell
//   h1[c][lane]             hopping from the centre to the slot-0 element of this lane (0 if they are not neighbours)
struct mom_args {
    const int32_t* f;
    int N, L, M, G, ncls;
    double U, mu_c, beta;
    double slot_val[FKMC_MAX_Z];
    const int* tabi;
    const int* off;
    const unsigned short* nb;
    const double* h1;
    const double* chebt;    // [M][G]
    const double* lobatto;  // [G]
    const double* dtheta;   // [G-1]
    double* moments;        // [B][M]
    double* ab;             // [B][4] e_min, e_max in; a, b out
    double* logz;           // [B]
    double* part;           // [B][MOM_SPLIT][3][FKMC_MAX_HALF + 1] partial traces of the CTAs of one proposal
    int* arrived;           // [B] CTAs of the proposal that have delivered their partial traces (self-resetting)
    // local re-evaluation (chain engine): f_cur = the chain's current configuration, ks_in = trace sums of f_cur under the scaling (a0, b0)
    // they were computed with, ks_out = the same record for f (written always when given), hop0 = hop distance from site 0
    const int32_t* f_cur;
    const double* ks_in;
    double* ks_out;
    const unsigned char* hop0;
    double guard;
};
// per-chain record of the local scheme: [3][FKMC_MAX_HALF + 1] sums (T_m)_jj, <T_(m-1) e_j|T_m e_j>, <T_m e_j|T_m e_j> over all columns j,
// then sum x, a0, b0, valid
// then sum x, a0, b0, valid, and what the sums were computed for: M, U, mu_c
constexpr int KS_TRX = 3 * (FKMC_MAX_HALF + 1), KS_A = KS_TRX + 1, KS_B = KS_TRX + 2, KS_VALID = KS_TRX + 3, KS_M = KS_TRX + 4, KS_U = KS_TRX + 5,
              KS_MU = KS_TRX + 6;
static_assert(KS_MU < FKMC_KPM_STATE, "state record too small");

constexpr int MOM_WARPS = 8;
constexpr int MOM_SPLIT = 2;  // CTAs per proposal (the columns are dealt out round robin): 2 B CTAs fill the last wave much better than B

// SCHED != 0: the slots per step are compile-time constants (four bits per step, m = 2 first), which turns the slot loop of a step
// into straight-line code whose loads and FMA chains interleave; UNI: all hoppings are equal (one multiplication per element).
template <int HALF, int Z, int S, unsigned long long SCHED, bool UNI>
__global__ void __launch_bounds__(MOM_WARPS * 32, 2) kpm_moments2d_kernel(mom_args P) {
    extern __shared__ __align__(16) double sm[];
    constexpr int EP = 32 * S, NT = HALF + 4;
    const int N = P.N, L = P.L, M = P.M, G = P.G;
    const int b = blockIdx.x / MOM_SPLIT, cta_part = blockIdx.x % MOM_SPLIT;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int PW = P.tabi[HALF + 2], PV = P.tabi[HALF + 3];
    const int PX = L + (((PW - L) % 16) + 16) % 16;  // row pitch of the diagonal: same residue structure as the patch array
    double* xd = sm;                           // [L][PX] diagonal of X
    double* red = xd + L * PX;                 // [48]
    double* msc = red + 48;                    // [64]
    double* Fg = msc + 64;                     // [G]
    double* acc = Fg + ((G + 1) & ~1);         // [MOM_WARPS][3][HALF+1]
    double* vec = acc + MOM_WARPS * 3 * (HALF + 1);  // [MOM_WARPS][2][PV]
    int* cols = reinterpret_cast<int*>(vec + (size_t)MOM_WARPS * 2 * PV);  // [N] columns of the local scheme

    const double e_min = P.ab[(size_t)b * 4 + 0], e_max = P.ab[(size_t)b * 4 + 1];
    const double a_new = (e_max - e_min) / 2., b_new = (e_max + e_min) / 2.;
    const int32_t* f = P.f + (size_t)b * N;
    // ---- local scheme?  The proposal differs from the chain's current configuration in one or two sites: only the columns within HALF
    // hops of those sites change, PROVIDED the scaling is kept; so the recursion runs under the scaling (a0, b0) of the stored sums for the
    // affected columns of both configurations, and the moments are re-expanded for the proposal's own scaling at the end.
    __shared__ int s_nd, s_sites[2], s_ncols, s_local;
    const int32_t* fc = P.f_cur ? P.f_cur + (size_t)b * N : nullptr;
    const double* ksi = P.ks_in ? P.ks_in + (size_t)b * FKMC_KPM_STATE : nullptr;
    if (tid == 0) s_nd = 0;
    __syncthreads();
    if (fc && ksi && P.ncls == 1) {
        for (int i = tid; i < N; i += T)
            if (f[i] != fc[i]) {
                const int k = atomicAdd(&s_nd, 1);
                if (k < 2) s_sites[k] = i;
            }
    }
    __syncthreads();
    if (tid == 0) {
        int loc = 0;
        if (fc && ksi && P.ncls == 1 && s_nd >= 1 && s_nd <= 2 && ksi[KS_VALID] == 1.0 && ksi[KS_M] == (double)M && ksi[KS_U] == P.U &&
            ksi[KS_MU] == P.mu_c) {
            const double alpha = ksi[KS_A] / a_new, beta_s = (ksi[KS_B] - b_new) / a_new;
            loc = fabs(alpha - 1.0) <= P.guard && fabs(beta_s) <= P.guard;
        }
        s_local = loc;
    }
    __syncthreads();
    const bool local = s_local != 0;
    const double a = local ? ksi[KS_A] : a_new, bsh = local ? ksi[KS_B] : b_new;   // scaling of the recursion
    double part = 0.0;
    for (int i = tid; i < N; i += T) {
        const double x = ((P.U * (double)f[i] - P.mu_c) - bsh) / a;
        const int y = i / L;
        xd[y * PX + (i - y * L)] = x;
        part += x;
    }
    const double trx = block_sum(part, red);
    int ncols = N;
    if (local) {
        // affected columns in increasing order (the order fixes which warp sums what: results do not depend on scheduling).
        // T_m e_j sees the diagonal entry of site i only if a walk from j reaches i in at most m - 1 <= HALF - 1 hops.
        __shared__ unsigned s_mask[64];
        __shared__ int s_sidx[64];
        const int nd = s_nd, nchunk = (N + 31) / 32;   // nchunk <= 64 (N <= 2048 on this path)
        for (int ch = warp; ch < nchunk; ch += MOM_WARPS) {
            const int j = ch * 32 + lane;
            bool hit = false;
            if (j < N) {
                const int yj = j / L, xj = j - yj * L;
                for (int k = 0; k < nd; ++k) {
                    const int sk = s_sites[k], ys = sk / L, xs = sk - ys * L;
                    int dy = yj - ys, dx = xj - xs;
                    dy += dy < 0 ? L : 0;
                    dx += dx < 0 ? L : 0;
                    hit = hit || P.hop0[dy * L + dx] < HALF;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_mask[ch] = m;
        }
        __syncthreads();
        if (warp == 0) {
            // exclusive prefix of the chunk counts (two chunks per lane)
            const int c0 = (2 * lane < nchunk) ? __popc(s_mask[2 * lane]) : 0, c1 = (2 * lane + 1 < nchunk) ? __popc(s_mask[2 * lane + 1]) : 0;
            int incl = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = incl - (c0 + c1);
            if (2 * lane < nchunk) s_sidx[2 * lane] = excl;
            if (2 * lane + 1 < nchunk) s_sidx[2 * lane + 1] = excl + c0;
            if (lane == 31) s_ncols = incl;
        }
        __syncthreads();
        for (int ch = warp; ch < nchunk; ch += MOM_WARPS) {
            const unsigned m = s_mask[ch];
            if (m >> lane & 1u) cols[s_sidx[ch] + __popc(m & ((1u << lane) - 1u))] = ch * 32 + lane;
        }
        __syncthreads();
        ncols = s_ncols;
    }
    double sv2[Z];
#pragma unroll
    for (int z = 0; z < Z; ++z) sv2[z] = 2.0 * (P.slot_val[z] / a);

    double* const s0 = vec + (size_t)warp * 2 * PV;  // T_even
    double* const s1 = s0 + PV;                      // T_odd
    for (int i = lane; i < 2 * PV; i += 32) s0[i] = 0.0;  // the cells outside the patch stay zero for good
    __syncwarp();

    double tr[HALF + 1], d01[HALF + 1], d11[HALF + 1];
#pragma unroll
    for (int m = 0; m <= HALF; ++m) tr[m] = d01[m] = d11[m] = 0.0;

    for (int pass = 0; pass < (local ? 2 : 1); ++pass) {
    if (pass == 1) {
        // second pass of the local scheme: the same columns for the current configuration, subtracted
        __syncthreads();
        if (tid < s_nd) {
            const int i = s_sites[tid], y = i / L;
            xd[y * PX + (i - y * L)] = ((P.U * (double)fc[i] - P.mu_c) - bsh) / a;
        }
#pragma unroll
        for (int m = 0; m <= HALF; ++m) { tr[m] = -tr[m]; d01[m] = -d01[m]; d11[m] = -d11[m]; }
        __syncthreads();
    }
    for (int cls = 0; cls < P.ncls; ++cls) {
        // this lane's elements: table index 32 s + lane
        int dyx[S], nbi[Z + 1][S], ns[HALF + 1];
#pragma unroll
        for (int m = 0; m <= HALF; ++m) ns[m] = P.tabi[cls * NT + m];
        const int lc = P.tabi[cls * NT + HALF + 1];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int e = 32 * s + lane;
            dyx[s] = P.off[cls * EP + e];
#pragma unroll
            for (int z = 0; z <= Z; ++z) nbi[z][s] = P.nb[((size_t)cls * (Z + 1) + z) * EP + e];
        }
        const double h1 = P.h1[cls * 32 + lane] / a;  // (X e_j)(e) for the slot-0 element next to the centre
        for (int idx = warp + MOM_WARPS * cta_part; idx < ncols; idx += MOM_WARPS * MOM_SPLIT) {
            const int j = local ? cols[idx] : idx;
            const int y0 = j / L, x0 = j - y0 * L;
            if (P.ncls > 1 && ((y0 + x0) & 1) != cls) continue;
            double xd2[S], va[S], vb[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                int yy = y0 + ((dyx[s] >> 8) & 255) - 128, xx = x0 + (dyx[s] & 255) - 128;
                yy += (yy < 0) ? L : 0; yy -= (yy >= L) ? L : 0;
                xx += (xx < 0) ? L : 0; xx -= (xx >= L) ? L : 0;
                xd2[s] = ((dyx[s] >> 16) <= HALF) ? 2.0 * xd[yy * PX + xx] : 0.0;
                va[s] = 0.0;
                vb[s] = 0.0;
                s0[nbi[Z][s]] = 0.0;   // what the previous column left in the patch
                s1[nbi[Z][s]] = 0.0;
            }
            // T_0 e_j = e_j (the centre), T_1 e_j = X e_j
            va[0] = (lane == lc) ? 1.0 : 0.0;
            vb[0] = (lane == lc) ? 0.5 * xd2[0] : h1;
            s1[nbi[Z][0]] = vb[0];
            __syncwarp();
            // one recursion step: vold <- 2 X vcur - vold, published to snew
            auto step = [&](double (&vold)[S], const double (&vcur)[S], const double* __restrict__ scur, double* __restrict__ snew,
                            const int m, const bool need_dots, double& q01, double& q11) {
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int nsm = SCHED ? (int)((SCHED >> (4 * (m - 2))) & 15) : ns[m];
                    if (s < nsm) {  // warp-uniform
                        double vn = fma(xd2[s], vcur[s], -vold[s]);
                        if (UNI) {
                            double nbs[Z];
#pragma unroll
                            for (int z = 0; z < Z; ++z) nbs[z] = scur[nbi[z][s]];
#pragma unroll
                            for (int w = 1; w < Z; w *= 2)
#pragma unroll
                                for (int z = 0; z + w < Z; z += 2 * w) nbs[z] += nbs[z + w];
                            vn = fma(sv2[0], nbs[0], vn);
                        } else {
#pragma unroll
                            for (int z = 0; z < Z; ++z) vn = fma(sv2[z], scur[nbi[z][s]], vn);
                        }
                        vn = ((dyx[s] >> 16) <= m) ? vn : 0.0;
                        vold[s] = vn;
                        snew[nbi[Z][s]] = vn;
                        if (need_dots) {
                            q01 = fma(vcur[s], vn, q01);
                            q11 = fma(vn, vn, q11);
                        }
                    }
                }
                __syncwarp();
            };
#pragma unroll
            for (int m = 2; m <= HALF; ++m) {
                const bool need_dots = (2 * m - 1 >= HALF);
                if ((m & 1) == 0) {
                    step(va, vb, s1, s0, m, need_dots, d01[m], d11[m]);  // va <- T_m e_j
                    tr[m] += (lane == lc) ? va[0] : 0.0;
                } else {
                    step(vb, va, s0, s1, m, need_dots, d01[m], d11[m]);  // vb <- T_m e_j
                    tr[m] += (lane == lc) ? vb[0] : 0.0;
                }
            }
        }
    }
    }
    if (local) {
#pragma unroll
        for (int m = 0; m <= HALF; ++m) { tr[m] = -tr[m]; d01[m] = -d01[m]; d11[m] = -d11[m]; }   // proposal minus current
    }
#pragma unroll
    for (int m = 2; m <= HALF; ++m) {
        const double a0 = warp_sum(tr[m]), a1 = warp_sum(d01[m]), a2 = warp_sum(d11[m]);
        if (lane == 0) {
            acc[(warp * 3 + 0) * (HALF + 1) + m] = a0;
            acc[(warp * 3 + 1) * (HALF + 1) + m] = a1;
            acc[(warp * 3 + 2) * (HALF + 1) + m] = a2;
        }
    }
    __syncthreads();
    // this CTA's traces (warps summed in a fixed order) -> global; the CTA that delivers last finishes the proposal
    if (MOM_SPLIT > 1) {
        double* mine = P.part + ((size_t)b * MOM_SPLIT + cta_part) * 3 * (FKMC_MAX_HALF + 1);
        if (tid < 3 * (HALF + 1)) {
            const int q = tid / (HALF + 1), m = tid % (HALF + 1);
            double t = 0.0;
            for (int w = 0; w < MOM_WARPS; ++w) t += acc[(w * 3 + q) * (HALF + 1) + m];
            __stcg(mine + q * (FKMC_MAX_HALF + 1) + m, t);
        }
        __threadfence();
        __syncthreads();
        __shared__ int last_flag;
        if (tid == 0) {
            const int seen = atomicAdd(P.arrived + b, 1);
            last_flag = (seen == MOM_SPLIT - 1);
            if (last_flag) P.arrived[b] = 0;  // ready for the next launch
        }
        __syncthreads();
        if (!last_flag) return;
        __threadfence();
    }
    // coefficients and logZ (configuration.cpp:198-202, chebyshev.hpp:36-54)
    double* mom = msc + 8;  // [M] (M <= 32)
    if (tid == 0) {
        bool is_set[2 * FKMC_MAX_HALF];
        for (int m = 0; m < M; ++m) { is_set[m] = false; mom[m] = 0.0; }
        mom[0] = 1.0; is_set[0] = true;
        mom[1] = trx / N; is_set[1] = true;
        for (int m = 2; m <= HALF; ++m) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            if (MOM_SPLIT > 1) {
                for (int c = 0; c < MOM_SPLIT; ++c) {
                    const double* pc = P.part + ((size_t)b * MOM_SPLIT + c) * 3 * (FKMC_MAX_HALF + 1);
                    t0 += __ldcg(pc + m);
                    t1 += __ldcg(pc + (FKMC_MAX_HALF + 1) + m);
                    t2 += __ldcg(pc + 2 * (FKMC_MAX_HALF + 1) + m);
                }
            } else {
                for (int w = 0; w < MOM_WARPS; ++w) {
                    t0 += acc[(w * 3 + 0) * (HALF + 1) + m];
                    t1 += acc[(w * 3 + 1) * (HALF + 1) + m];
                    t2 += acc[(w * 3 + 2) * (HALF + 1) + m];
                }
            }
            if (local) {   // sums of the proposal = sums of the current configuration + (affected columns: proposal - current)
                t0 += ksi[m];
                t1 += ksi[(FKMC_MAX_HALF + 1) + m];
                t2 += ksi[2 * (FKMC_MAX_HALF + 1) + m];
            }
            if (P.ks_out) {
                double* kso = P.ks_out + (size_t)b * FKMC_KPM_STATE;
                kso[m] = t0;
                kso[(FKMC_MAX_HALF + 1) + m] = t1;
                kso[2 * (FKMC_MAX_HALF + 1) + m] = t2;
            }
            if (!is_set[m]) { mom[m] = t0 / N; is_set[m] = true; }
            int kk = 2 * m - 1;
            if (kk < M && kk >= HALF) {
                mom[kk] = (t1 * 2. - trx) / N; is_set[kk] = true;
                if (kk != M - 1) { ++kk; mom[kk] = (t2 / N * 2. - 1.0); is_set[kk] = true; }
            }
        }
    }
    if (tid == 0 && P.ks_out) {
        double* kso = P.ks_out + (size_t)b * FKMC_KPM_STATE;
        kso[KS_TRX] = trx; kso[KS_A] = a; kso[KS_B] = bsh; kso[KS_VALID] = 1.0;
        kso[KS_M] = (double)M; kso[KS_U] = P.U; kso[KS_MU] = P.mu_c;
    }
    if (local) {
        // moments under the proposal's own scaling y = (H - b_new) / a_new = alpha x + beta with x = (H - b0) / a0:
        // T_m(alpha x + beta) = sum_k c_mk T_k(x), rows by the three-term recurrence in the Chebyshev basis (x T_k = (T_(k+1) + T_|k-1|) / 2);
        // lane k of warp 0 holds coefficient k of the two latest rows, the products c_mk mu_k go to shared memory and are summed per row
        __syncthreads();   // mom[] of thread 0
        double* cm = vec;  // [M][32] (the patch buffers are free now)
        if (warp == 0) {
            const double alpha = a / a_new, beta_s = (bsh - b_new) / a_new;
            const double mk = (lane < M) ? mom[lane] : 0.0;
            double pm1 = (lane == 0) ? 1.0 : 0.0;                                   // P_0
            double pm = (lane == 0) ? beta_s : ((lane == 1) ? alpha : 0.0);         // P_1
            cm[0 * 32 + lane] = pm1 * mk;
            cm[1 * 32 + lane] = pm * mk;
            for (int m = 2; m < M; ++m) {
                const double up = __shfl_down_sync(0xffffffffu, pm, 1);              // p_(k+1)
                double dn = __shfl_up_sync(0xffffffffu, pm, 1);                      // p_(k-1)
                dn = (lane == 0) ? 0.0 : ((lane == 1) ? 2.0 * dn : dn);
                const double pn = (lane <= m) ? fma(alpha, (lane == 31 ? 0.0 : up) + dn, fma(2.0 * beta_s, pm, -pm1)) : 0.0;
                cm[m * 32 + lane] = pn * mk;
                pm1 = pm;
                pm = pn;
            }
        }
        __syncthreads();
        if (tid < M) {
            double sacc = 0.0;
            for (int k = M - 1; k >= 0; --k) sacc += cm[tid * 32 + k];
            mom[tid] = sacc;
        }
        __syncthreads();
    }
    for (int i = tid; i < G; i += T) Fg[i] = N * log(1. + exp(-P.beta * (a_new * P.lobatto[i] + b_new)));
    __syncthreads();
    if (tid < M) {
        const double* Tm = P.chebt + (size_t)tid * G;
        double s = 0.0;
        for (int i = 0; i < G - 1; ++i) s += (Fg[i + 1] * Tm[i + 1] + Fg[i] * Tm[i]) * P.dtheta[i];
        acc[tid] = s * 0.5;
    }
    __syncthreads();
    if (tid == 0) {
        double s = acc[0];
        for (int m = 1; m < M; ++m) s += 2. * acc[m] * mom[m];
        P.logz[b] = s;
        P.ab[(size_t)b * 4 + 2] = a_new;
        P.ab[(size_t)b * 4 + 3] = b_new;
    }
    if (tid < M && P.moments) P.moments[(size_t)b * M + tid] = mom[tid];
}

// number of sites within H hops on the infinite lattice (upper bound for the honeycomb brick wall)
constexpr int patch_elems(int kind, int H) { return kind == FKMC_TRIANGULAR ? 3 * H * H + 3 * H + 1 : 2 * H * H + 2 * H + 1; }

}  // namespace

// Build (or reuse) the patch tables for radius H (see the comment above mom_args).  *S_out = slots per lane.
static int prepare_patch_tables(fkmc_ctx* ctx, int H, int* S_out) {
    if (ctx->kpm2_H == H && ctx->d_kpm2_tabi) { *S_out = ctx->kpm2_S; return FKMC_OK; }
    const int N = ctx->N, L = ctx->L, Z = ctx->Z, NT = H + 4;
    const int ncls = (ctx->kind == FKMC_HONEYCOMB) ? 2 : 1;
    struct site { int s, dy, dx, hops; };
    std::vector<std::vector<site>> patch(ncls);
    std::vector<int> centre(ncls);
    for (int c = 0; c < ncls; ++c) {
        // the centre sits in the middle of the lattice and L >= 2H+1, so the patch never crosses the periodic boundary
        const int y0 = L / 2, x0 = ((y0 + L / 2) & 1) == c ? L / 2 : L / 2 - 1;
        const int j0 = y0 * L + x0;
        centre[c] = j0;
        std::vector<int> dist(N, -1);
        std::queue<int> q;
        dist[j0] = 0;
        q.push(j0);
        while (!q.empty()) {
            const int s = q.front();
            q.pop();
            patch[c].push_back({s, s / L - y0, s % L - x0, dist[s]});
            if (dist[s] == H) continue;
            for (int z = 0; z < Z; ++z) {
                const int t = ctx->h_nbr_idx[(size_t)z * N + s];
                if (t < N && dist[t] < 0) { dist[t] = dist[s] + 1; q.push(t); }
            }
        }
        for (auto& e : patch[c])
            if (std::abs(e.dy) > H || std::abs(e.dx) > H) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: patch wraps around the lattice");
        std::stable_sort(patch[c].begin(), patch[c].end(), [](const site& a, const site& b2) {
            if (a.hops != b2.hops) return a.hops < b2.hops;
            return std::atan2((double)a.dy, (double)a.dx) < std::atan2((double)b2.dy, (double)b2.dx);
        });
    }
    // deal the elements out, nearest first: an element goes to the first slot whose half-warps still have its lane
    // (= cell address mod 16) free.  The row pitch P decides the residues; take the one that needs the fewest slot visits.
    auto cell = [&](int P, int dy, int dx) { return (dy + H + 1) * P + (dx + H + 1); };
    struct plan { long visits; int S; std::vector<std::vector<int>> where; std::vector<std::vector<int>> ns; };  // where[c][e] = 32 slot + lane
    auto make_plan = [&](int P) {
        plan pl;
        pl.visits = 0;
        pl.S = 0;
        pl.where.resize(ncls);
        pl.ns.assign(ncls, std::vector<int>(H + 1, 1));
        for (int c = 0; c < ncls; ++c) {
            std::vector<char> used;  // [slot][half][residue]
            for (const site& e : patch[c]) {
                const int r = cell(P, e.dy, e.dx) & 15;
                int sl = 0, half = -1;
                for (;; ++sl) {
                    if ((int)used.size() < 32 * (sl + 1)) used.resize(32 * (sl + 1), 0);
                    if (!used[32 * sl + r]) { half = 0; break; }
                    if (!used[32 * sl + 16 + r]) { half = 1; break; }
                }
                used[32 * sl + 16 * half + r] = 1;
                pl.where[c].push_back(32 * sl + 16 * half + r);
                for (int m = e.hops; m <= H; ++m) pl.ns[c][m] = std::max(pl.ns[c][m], sl + 1);
                pl.S = std::max(pl.S, sl + 1);
            }
            for (int m = 2; m <= H; ++m) pl.visits += pl.ns[c][m];
        }
        return pl;
    };
    int P = 2 * H + 3;
    plan best = make_plan(P);
    for (int cand = 2 * H + 4; cand < 2 * H + 3 + 16; ++cand) {
        plan pl = make_plan(cand);
        if (pl.visits < best.visits || (pl.visits == best.visits && pl.S < best.S)) { best = pl; P = cand; }
    }
    const int S = best.S, EP = 32 * S;
    int PV = (2 * H + 3) * P;
    PV += PV & 1;
    if (PV > 65000) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: patch larger than its table");
    std::vector<int> tabi((size_t)ncls * NT, 0), off((size_t)ncls * EP, (255 << 16) | (128 << 8) | 128);
    std::vector<unsigned short> nb((size_t)ncls * (Z + 1) * EP);
    std::vector<double> h1((size_t)ncls * 32, 0.0);
    for (int c = 0; c < ncls; ++c) {
        for (int m = 0; m <= H; ++m) tabi[(size_t)c * NT + m] = best.ns[c][m];
        tabi[(size_t)c * NT + H + 2] = P;
        tabi[(size_t)c * NT + H + 3] = PV;
        // unused (slot, lane): a cell of the first row (outside the patch) with the lane's residue, so that it stays conflict-free
        for (int e = 0; e < EP; ++e)
            for (int z = 0; z <= Z; ++z) nb[((size_t)c * (Z + 1) + z) * EP + e] = (unsigned short)(e & 15);
        for (size_t i = 0; i < patch[c].size(); ++i) {
            const site& e = patch[c][i];
            const int w = best.where[c][i];
            off[(size_t)c * EP + w] = (e.hops << 16) | ((e.dy + 128) << 8) | (e.dx + 128);
            nb[((size_t)c * (Z + 1) + Z) * EP + w] = (unsigned short)cell(P, e.dy, e.dx);
            for (int z = 0; z < Z; ++z) {
                const int t = ctx->h_nbr_idx[(size_t)z * N + e.s];
                // a padding slot of the stencil (no neighbour): any cell that stays zero, keep the residue
                int cw = cell(P, e.dy, e.dx) & 15;
                if (t < N && t != e.s) {
                    int ty = t / L - centre[c] / L, tx = t % L - centre[c] % L;  // minimal image: a neighbour just outside the patch may wrap
                    ty += (ty < -(H + 1)) ? L : 0; ty -= (ty > H + 1) ? L : 0;
                    tx += (tx < -(H + 1)) ? L : 0; tx -= (tx > H + 1) ? L : 0;
                    if (std::abs(ty) > H + 1 || std::abs(tx) > H + 1) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: neighbour outside the patch array");
                    cw = cell(P, ty, tx);
                }
                nb[((size_t)c * (Z + 1) + z) * EP + w] = (unsigned short)cw;
                if (t == centre[c] && w < 32) h1[(size_t)c * 32 + w] += ctx->h_nbr_val[(size_t)z * N + e.s];
            }
            if (e.hops == 0) tabi[(size_t)c * NT + H + 1] = w;
            if (e.hops <= 1 && w >= 32) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: first ring does not fit the first slot");
        }
        if (getenv("FKMC_KPM_DEBUG")) {
            fprintf(stderr, "kpm2d patch tables: class %d, %zu elements, pitch %d, %d slots, slots per step:", c, patch[c].size(), P, S);
            for (int m = 2; m <= H; ++m) fprintf(stderr, " %d", best.ns[c][m]);
            fprintf(stderr, "\n");
        }
    }
    if (ctx->d_kpm2_tabi) {
        cudaFree(ctx->d_kpm2_tabi); cudaFree(ctx->d_kpm2_off); cudaFree(ctx->d_kpm2_nb); cudaFree(ctx->d_kpm2_h1);
        ctx->d_kpm2_tabi = nullptr;
    }
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_kpm2_tabi, sizeof(int) * tabi.size()));
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_kpm2_off, sizeof(int) * off.size()));
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_kpm2_nb, sizeof(unsigned short) * nb.size()));
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_kpm2_h1, sizeof(double) * h1.size()));
    FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_kpm2_tabi, tabi.data(), sizeof(int) * tabi.size(), cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_kpm2_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_kpm2_nb, nb.data(), sizeof(unsigned short) * nb.size(), cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_kpm2_h1, h1.data(), sizeof(double) * h1.size(), cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->kpm2_sched = 0;
    for (int m = 2; m <= H && m <= 17; ++m) ctx->kpm2_sched |= (unsigned long long)(best.ns[0][m] & 15) << (4 * (m - 2));
    for (int c = 1; c < ncls; ++c)
        for (int m = 2; m <= H; ++m)
            if (best.ns[c][m] != best.ns[0][m]) ctx->kpm2_sched = ~0ull;
    ctx->kpm2_H = H;
    ctx->kpm2_S = S;
    ctx->kpm2_PV = PV;
    ctx->kpm2_P = P;
    *S_out = S;
    return FKMC_OK;
}

template <int KIND>
static int launch_lanczos2d(fkmc_ctx* ctx, const lz_args& P, int B) {
    const size_t smem = sizeof(double) * (2 * (size_t)P.N + 2 * (LZ_KMAX + 2) + 2 * (LZ_T / 32) + 8);
    FKMC_CUDA(ctx, cudaFuncSetAttribute(lanczos2d_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lanczos2d_kernel<KIND><<<B, LZ_T, smem, ctx->stream>>>(P);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

template <int HALF, int Z, int S, unsigned long long SCHED, bool UNI>
static int launch_moments2d_s(fkmc_ctx* ctx, mom_args& P, int B) {
    const int PW = ctx->kpm2_P, PX = P.L + (((PW - P.L) % 16) + 16) % 16;
    const size_t smem = sizeof(double) * ((size_t)P.L * PX + 48 + 64 + ((P.G + 1) & ~1) + (size_t)MOM_WARPS * 3 * (HALF + 1) +
                                          (size_t)MOM_WARPS * 2 * ctx->kpm2_PV) + sizeof(int) * (size_t)((P.N + 1) & ~1);
    if (smem > ctx->smem_optin - 1024) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: lattice too large for the shared-memory kernel");
    FKMC_CUDA(ctx, cudaFuncSetAttribute(kpm_moments2d_kernel<HALF, Z, S, SCHED, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kpm_moments2d_kernel<HALF, Z, S, SCHED, UNI><<<B * MOM_SPLIT, MOM_WARPS * 32, smem, ctx->stream>>>(P);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

// slots per step of the square lattice's diamond patch, m = 2, 3, ... (what prepare_patch_tables finds; checked at launch)
constexpr unsigned long long cubic2d_sched(int H) {
    constexpr int ns[9] = {1, 1, 2, 3, 3, 4, 5, 6, 8};
    unsigned long long v = 0;
    for (int m = 2; m <= H && m <= 10; ++m) v |= (unsigned long long)ns[m - 2] << (4 * (m - 2));
    return v;
}

template <int HALF, int KIND>
static int launch_moments2d(fkmc_ctx* ctx, mom_args& P, int B) {
    constexpr int Z = (KIND == FKMC_TRIANGULAR) ? 6 : (KIND == FKMC_HONEYCOMB ? 3 : 4);
    constexpr int S0 = (patch_elems(KIND, HALF) + 31) / 32;  // slots if every slot could be filled completely
    int S = 0;
    int rc = prepare_patch_tables(ctx, HALF, &S);
    if (rc) return rc;
    P.tabi = ctx->d_kpm2_tabi;
    P.off = ctx->d_kpm2_off;
    P.nb = ctx->d_kpm2_nb;
    P.h1 = ctx->d_kpm2_h1;
    if (!ctx->d_kpm2_part) {
        FKMC_CUDA(ctx, cudaMalloc(&ctx->d_kpm2_part, sizeof(double) * (size_t)ctx->max_batch * MOM_SPLIT * 3 * (FKMC_MAX_HALF + 1)));
        FKMC_CUDA(ctx, cudaMalloc(&ctx->d_kpm2_arrived, sizeof(int) * (size_t)ctx->max_batch));
        FKMC_CUDA(ctx, cudaMemsetAsync(ctx->d_kpm2_arrived, 0, sizeof(int) * (size_t)ctx->max_batch, ctx->stream));
    }
    P.part = ctx->d_kpm2_part;
    P.arrived = ctx->d_kpm2_arrived;
    if constexpr (KIND == FKMC_CUBIC2D) {
        // the common case, fully specialised: compile-time schedule, one hopping constant
        bool uni = true;
        for (int z = 1; z < Z; ++z) uni = uni && P.slot_val[z] == P.slot_val[0];
        constexpr int SS = S0 + (HALF == 10 ? 1 : 0);
        if (uni && S == SS && ctx->kpm2_sched == cubic2d_sched(HALF) && !ctx->kpm_no_sched)
            return launch_moments2d_s<HALF, Z, SS, cubic2d_sched(HALF), true>(ctx, P, B);
    }
    // the honeycomb patch is smaller than the bound used for S0, and the residue constraint can cost one slot
    if (S <= S0 - 1 && S0 >= 2) return launch_moments2d_s<HALF, Z, (S0 >= 2 ? S0 - 1 : 1), 0ull, false>(ctx, P, B);
    if (S <= S0) return launch_moments2d_s<HALF, Z, S0, 0ull, false>(ctx, P, B);
    if (S == S0 + 1) return launch_moments2d_s<HALF, Z, S0 + 1, 0ull, false>(ctx, P, B);
    return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: patch does not fit the compiled slot counts");
}

template <int KIND>
static int launch_moments2d_half(fkmc_ctx* ctx, mom_args& P, int B, int half) {
    switch (half) {
        case 2: return launch_moments2d<2, KIND>(ctx, P, B);
        case 3: return launch_moments2d<3, KIND>(ctx, P, B);
        case 4: return launch_moments2d<4, KIND>(ctx, P, B);
        case 5: return launch_moments2d<5, KIND>(ctx, P, B);
        case 6: return launch_moments2d<6, KIND>(ctx, P, B);
        case 7: return launch_moments2d<7, KIND>(ctx, P, B);
        case 8: return launch_moments2d<8, KIND>(ctx, P, B);
        case 9: return launch_moments2d<9, KIND>(ctx, P, B);
        case 10: return launch_moments2d<10, KIND>(ctx, P, B);
    }
    return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: unsupported M for the 2-D kernels");
}

// Hop distances from site 0 (breadth-first search over the stencil): what the local scheme needs to list the columns near a changed site.
// Called outside any stream capture (chain set-up, fkmc_logz_kpm_batched_local).
int fkmc_kpm_prepare_local(fkmc_ctx* ctx) {
    if (ctx->d_kpm_hop0) return FKMC_OK;
    const int N = ctx->N, Z = ctx->Z;
    std::vector<unsigned char> hop(N, 255);
    std::queue<int> q;
    hop[0] = 0;
    q.push(0);
    while (!q.empty()) {
        const int sidx = q.front();
        q.pop();
        for (int z = 0; z < Z; ++z) {
            const int t = ctx->h_nbr_idx[(size_t)z * N + sidx];
            if (t < N && ctx->h_nbr_val[(size_t)z * N + sidx] != 0.0 && hop[t] == 255 && hop[sidx] < 254) { hop[t] = hop[sidx] + 1; q.push(t); }
        }
    }
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_kpm_hop0, N));
    FKMC_CUDA(ctx, cudaMemcpy(ctx->d_kpm_hop0, hop.data(), N, cudaMemcpyHostToDevice));
    return FKMC_OK;
}

// true when (lattice, M) is served by the two-kernel path
bool fkmc_kpm2d_applicable(const fkmc_ctx* ctx, int M) {
    const int half = M / 2;
    if (ctx->kpm_force_generic || ctx->kpm_force_v1) return false;
    if (ctx->kind != FKMC_CUBIC2D && ctx->kind != FKMC_TRIANGULAR && ctx->kind != FKMC_HONEYCOMB) return false;
    if (ctx->L % 8 != 0 || ctx->N > 8 * LZ_T) return false;
    return half >= 2 && half <= 10 && ctx->L >= 2 * half + 1;
}

int fkmc_launch_kpm2d(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, int M, int G, const double* slot_val,
                      double* d_moments, double* d_ab, double* d_logz) {
    lz_args Q{};
    Q.f = d_f; Q.N = ctx->N; Q.L = ctx->L; Q.U = U; Q.mu_c = mu_c;
    Q.ht = slot_val[0];
    Q.hp = (ctx->kind == FKMC_TRIANGULAR) ? slot_val[4] : 0.0;
    Q.ab = d_ab; Q.flag = ctx->d_flag; Q.steps = ctx->d_kpm_steps;
    Q.kmax = ctx->lanczos_cap > 0 ? std::min(ctx->lanczos_cap, LZ_KMAX) : LZ_KMAX;
    if (!ctx->d_kpm2_order) {
        FKMC_CUDA(ctx, cudaMalloc(&ctx->d_kpm2_order, sizeof(int) * (size_t)ctx->max_batch));
        FKMC_CUDA(ctx, cudaMemsetAsync(ctx->d_kpm_steps, 0, sizeof(int) * (size_t)ctx->max_batch, ctx->stream));
    }
    if (B >= 2 * ctx->num_sms) {  // below that every CTA starts at once anyway
        lanczos_order_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_kpm_steps, B, ctx->d_kpm2_order);
        ctx->launches++;
        Q.order = ctx->d_kpm2_order;
    }
    int rc;
    {
        fkmc_prof_scope ps(ctx, "kpm_lanczos");
        if (ctx->kind == FKMC_CUBIC2D) rc = launch_lanczos2d<FKMC_CUBIC2D>(ctx, Q, B);
        else if (ctx->kind == FKMC_TRIANGULAR) rc = launch_lanczos2d<FKMC_TRIANGULAR>(ctx, Q, B);
        else rc = launch_lanczos2d<FKMC_HONEYCOMB>(ctx, Q, B);
    }
    if (rc) return rc;
    mom_args P{};
    P.f = d_f; P.N = ctx->N; P.L = ctx->L; P.M = M; P.G = G; P.ncls = (ctx->kind == FKMC_HONEYCOMB) ? 2 : 1;
    P.U = U; P.mu_c = mu_c; P.beta = beta;
    for (int z = 0; z < FKMC_MAX_Z; ++z) P.slot_val[z] = slot_val[z];
    P.chebt = ctx->d_chebt; P.lobatto = ctx->d_lobatto; P.dtheta = ctx->d_dtheta;
    P.moments = d_moments; P.ab = d_ab; P.logz = d_logz;
    // local re-evaluation (set by the chain engine around its launches; translation-invariant one-class lattices only)
    P.ks_out = ctx->kpm_ks_out;
    ctx->kpm_state_written = P.ks_out != nullptr;
    P.guard = 0.02;
    if (ctx->kpm_f_cur && ctx->kpm_ks_in && P.ncls == 1 && ctx->kpm_local && ctx->d_kpm_hop0) {
        P.f_cur = ctx->kpm_f_cur; P.ks_in = ctx->kpm_ks_in; P.hop0 = ctx->d_kpm_hop0;
    }
    if (ctx->kpm_wait_event) {  // reference configurations / records uploaded on the copy stream (fkmc_logz_kpm_batched_local)
        FKMC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->kpm_wait_event, 0));
        ctx->kpm_wait_event = nullptr;
    }
    fkmc_prof_scope ps(ctx, "kpm_moments");
    if (ctx->kind == FKMC_CUBIC2D) return launch_moments2d_half<FKMC_CUBIC2D>(ctx, P, B, M / 2);
    if (ctx->kind == FKMC_TRIANGULAR) return launch_moments2d_half<FKMC_TRIANGULAR>(ctx, P, B, M / 2);
    return launch_moments2d_half<FKMC_HONEYCOMB>(ctx, P, B, M / 2);
}
