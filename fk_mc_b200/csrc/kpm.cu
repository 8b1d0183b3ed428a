// Batched Chebyshev / KPM weight evaluation: one CTA per proposed configuration.
//
// Replaces configuration_t::calc_chebyshev (src/configuration.cpp:94-205) and
// chebyshev_eval::{ctor, moment_f, moment} (include/fk_mc/chebyshev.hpp:21-54):
//   1. e_min / e_max of H (reference: two ARPACK solves, configuration.cpp:99-100) by a Lanczos
//      iteration held in shared memory, Ritz values by warp-wide multisection on the Lanczos
//      tridiagonal, stopped when both ends stagnate;
//   2. a = (e_max-e_min)/2, b = (e_max+e_min)/2, X = (H-b)/a; exact full-trace moments
//      mu_m = Tr T_m(X)/N by the column recursion T_m e_j = 2 X T_{m-1} e_j - T_{m-2} e_j, m <= M/2,
//      and the doubling identities for M/2 <= k < M (configuration.cpp:117-194).  The hopping is
//      applied as a shared-memory stencil (slot-major neighbour table, per-slot hopping constants);
//      each warp owns a set of columns j and keeps two iterates in shared memory, updating in place;
//   3. c_m = moment_f(N log(1+e^{-beta(a x+b)}), m) on the G-point Lobatto grid (trapezoid rule) and
//      logZ = c_0 + 2 sum_{m>=1} c_m mu_m (configuration.cpp:198-202).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>

#include "common.cuh"

namespace {

constexpr int KPM_KMAX = 384;  // Lanczos step cap

struct kpm_args {
    const int32_t* f;
    const int* nbr_idx;  // [Z][N] global
    int N, Z, M, G;
    double U, mu_c, beta;
    double slot_val[FKMC_MAX_Z];
    const double* chebt;    // [M][G]
    const double* lobatto;  // [G]
    const double* dtheta;   // [G-1]
    double* moments;        // [B][M]
    double* ab;             // [B][4]
    double* logz;           // [B]
    int* flag;
    int* steps;             // [B] Lanczos steps taken (diagnostics), may be null
    int L;                  // linear lattice size (patch variant)
    int nwarps_cols;        // warps that run the column recursion
    int vec_doubles;        // size of the shared vector region (>= 2 Nv per column warp and >= the Lanczos need)
    int kmax;               // Lanczos step cap (KPM_KMAX unless the "lanczos_max_steps" option lowers it)
};

// number of eigenvalues of the k x k Lanczos tridiagonal below x.  ab[i] = (alpha_i, beta_i^2) with beta_0 = 0
// (beta_i couples rows i-1 and i).  Rows are fetched eight at a time so that the shared-memory latency is paid once
// per chunk and only the dependent FMA chain remains.
__device__ __forceinline__ int lanczos_sturm(const double2* __restrict__ ab, int k, double x) {
    double pm1 = 1.0, p = ab[0].x - x;
    if (p == 0.0) p = -DBL_EPSILON;
    bool neg = p < 0.0;
    int cnt = neg ? 1 : 0;
    int i = 1;
    for (; i + 8 <= k; i += 8) {
        double2 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = ab[i + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double pn = fma(q[u].x - x, p, -(q[u].y * pm1));
            if (pn == 0.0) pn = -DBL_EPSILON * p;
            const bool nneg = pn < 0.0;
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
        double pn = fma(q.x - x, p, -(q.y * pm1));
        if (pn == 0.0) pn = -DBL_EPSILON * p;
        const bool nneg = pn < 0.0;
        cnt += (nneg != neg) ? 1 : 0;
        neg = nneg;
        pm1 = p;
        p = pn;
    }
    return cnt;
}

// idx-th eigenvalue of the Lanczos tridiagonal by 32-way multisection (whole warp participates)
__device__ __forceinline__ double warp_ritz(const double2* __restrict__ ab, int k, int idx, double lo, double hi, int lane) {
    const double pad = 8.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) + DBL_MIN;
    double a = lo - pad, c = hi + pad;
    for (int round = 0; round < 14; ++round) {
        const double h = (c - a) * (1.0 / 33.0);
        if (!(h > 2.0 * DBL_EPSILON * fmax(fabs(a), fabs(c)) * (1.0 / 33.0))) break;
        const double x = a + h * (double)(lane + 1);
        const bool above = lanczos_sturm(ab, k, x) > idx;
        const unsigned mask = __ballot_sync(0xffffffffu, above);
        const int first = mask ? (__ffs(mask) - 1) : 32;
        const double na = first > 0 ? a + h * (double)first : a;
        const double nc = first < 32 ? a + h * (double)(first + 1) : c;
        a = na;
        c = nc;
    }
    return 0.5 * (a + c);
}

// KIND = 0: generic stencil on full lattice vectors (any lattice, any L).
// KIND = FKMC_CUBIC2D / FKMC_TRIANGULAR / FKMC_HONEYCOMB: local-patch variant for 2-D lattices with L >= 2 HALF + 1.
// T_m(X) e_j is supported within m hops of site j, so the column recursion runs on the (2 HALF+1)^2 patch
// around j, flattened with row stride 2 HALF+1 and a zero guard zone; the stencil becomes fixed offsets
// (+-1, +-PW, +-(PW+1)) with no index table, and a warp needs 2 x 2.6 KB of shared memory instead of 2 x 8 KB.
template <int HALF, int KIND>
__global__ void __launch_bounds__(KIND ? 256 : 384, KIND ? 2 : 1) kpm_kernel(kpm_args P) {
    extern __shared__ double sm[];
    const int N = P.N, Z = P.Z, M = P.M, G = P.G;
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int Nv = N + 1;  // vectors carry a trailing zero slot (padding neighbours point at it)
    // ---- shared-memory carve-up ----
    double* xd = sm;                       // [N]   diagonal (later scaled)
    double* red = xd + N;                  // [48]
    double* msc = red + 48;                // [64]  scalars / moments / coefficients
    double* Fg = msc + 64;                 // [G]   F(x_i) on the Lobatto grid
    double* acc = Fg + ((G + 1) & ~1);     // [nwarps][3][HALF+1] per-warp partial traces
    double* vec = acc + nwarps * 3 * (HALF + 1);  // [nwarps_cols][2][Nv]
    unsigned short* nidx = reinterpret_cast<unsigned short*>(vec + P.vec_doubles);  // [Z][N]

    const int32_t* f = P.f + (size_t)b * N;
    for (int i = tid; i < N; i += T) xd[i] = P.U * (double)f[i] - P.mu_c;
    for (int i = tid; i < Z * N; i += T) nidx[i] = (unsigned short)P.nbr_idx[i];
    __syncthreads();

    // =========================== 1. Lanczos for e_min / e_max ===========================
    // v_k is kept twice: in registers (own elements, together with v_{k-1} and w) and in a ping-pong pair of shared
    // vectors for the neighbour reads of the stencil.  One fused reduction per step gives alpha = v.Hv' and |w|^2,
    // beta^2 = |w|^2 - alpha^2 (recomputed directly when it cancels), so a step costs two barriers.
    constexpr int LE = 8;                // own elements per thread (N <= 8 * blockDim)
    double* vb0 = vec;                   // v_k, ping
    double* vb1 = vec + Nv;              // v_k, pong
    double2* ab = reinterpret_cast<double2*>((reinterpret_cast<uintptr_t>(vec + 3 * Nv) + 15) & ~(uintptr_t)15);  // [KPM_KMAX + 1] (alpha_i, beta_i^2)
    double* redp = red;                  // [2][nwarps][2] ping-pong partial sums (nwarps <= 12 -> 48 doubles)
    double vr[LE], pr[LE], wr[LE];
    double hv[FKMC_MAX_Z];  // per-slot hopping constants in registers
#pragma unroll
    for (int z = 0; z < FKMC_MAX_Z; ++z) hv[z] = (z < Z) ? P.slot_val[z] : 0.0;
    {
        double part = 0.0;
#pragma unroll
        for (int q = 0; q < LE; ++q) {
            const int i = tid + q * T;
            double v = 0.0;
            if (i < N) {
                unsigned h = (unsigned)i * 2654435761u + 0x9e3779b9u;
                h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
                v = (double)h * (1.0 / 4294967296.0) - 0.5;
            }
            vr[q] = v; pr[q] = 0.0; wr[q] = 0.0;
            part = fma(v, v, part);
        }
        if (tid == 0) { vb0[N] = 0.0; vb1[N] = 0.0; ab[0].y = 0.0; }  // zero slots
        const double nrm = sqrt(block_sum(part, msc));
#pragma unroll
        for (int q = 0; q < LE; ++q) {
            const int i = tid + q * T;
            vr[q] /= nrm;
            if (i < N) vb0[i] = vr[q];
        }
        __syncthreads();
    }
#ifdef FKMC_KPM_TIMING
    long long t_ritz = 0, t_start = clock64(), t_tmp = 0;
#endif
    const int kcap = min(P.kmax, N);
    double e_min = 0.0, e_max = 0.0, gl = DBL_MAX, gh = -DBL_MAX, hscale = 0.0;
    bool converged = false;
    int k = 0;
    double beta_k = 0.0;
    while (k < kcap) {
        const double* lv = (k & 1) ? vb1 : vb0;
        double* lvn = (k & 1) ? vb0 : vb1;
        double* rp = redp + (k & 1) * 2 * nwarps;
        double pa = 0.0;
#pragma unroll
        for (int q = 0; q < LE; ++q) {
            const int i = tid + q * T;
            if (i < N) {
                // all neighbour indices first, then the gathers: the loads of different slots overlap
                int nb[FKMC_MAX_Z];
#pragma unroll
                for (int z = 0; z < FKMC_MAX_Z; ++z) nb[z] = (z < Z) ? nidx[z * N + i] : N;
                double sacc = xd[i] * vr[q];
#pragma unroll
                for (int z = 0; z < FKMC_MAX_Z; ++z)
                    if (z < Z) sacc = fma(hv[z], lv[nb[z]], sacc);
                sacc = fma(-beta_k, pr[q], sacc);
                wr[q] = sacc;
                pa = fma(sacc, vr[q], pa);
            }
        }
        pa = warp_sum(pa);
        if (lane == 0) rp[2 * warp] = pa;
        __syncthreads();
        double alpha = 0.0;
#pragma unroll
        for (int ww = 0; ww < 12; ++ww)
            if (ww < nwarps) alpha += rp[2 * ww];
        // classical two-reduction form (the one-reduction shortcut |w|^2 - alpha^2 lets the normalisation error grow
        // like (alpha/beta)^2 per step): w -= alpha v, then its norm
        double pb = 0.0;
#pragma unroll
        for (int q = 0; q < LE; ++q) {
            wr[q] = fma(-alpha, vr[q], wr[q]);
            pb = fma(wr[q], wr[q], pb);
        }
        pb = warp_sum(pb);
        if (lane == 0) rp[2 * warp + 1] = pb;
        __syncthreads();
        double nb2 = 0.0;
#pragma unroll
        for (int ww = 0; ww < 12; ++ww)
            if (ww < nwarps) nb2 += rp[2 * ww + 1];
        const double nb = sqrt(fmax(nb2, 0.0));
        if (tid == 0) { ab[k].x = alpha; ab[k + 1].y = nb * nb; }
        // Gershgorin enclosure of the Lanczos tridiagonal (all threads keep it in registers)
        gl = fmin(gl, alpha - beta_k - nb);
        gh = fmax(gh, alpha + beta_k + nb);
        hscale = fmax(hscale, fabs(alpha) + nb);
        ++k;
        const bool breakdown = nb <= 1e-13 * hscale;
        if (!breakdown) {
            const double inv = 1.0 / nb;
#pragma unroll
            for (int q = 0; q < LE; ++q) {
                const int i = tid + q * T;
                const double vn = wr[q] * inv;
                pr[q] = vr[q];
                vr[q] = vn;
                if (i < N) lvn[i] = vn;
            }
        }
        beta_k = nb;
        __syncthreads();
        const bool last = breakdown || k == kcap;
        if (last || (k >= 96 && (k & 31) == 0)) {
            // Ritz values of T_k and of T_{k-16} (four independent multisections spread over the warps); converged when
            // both ends have stopped moving.  beta_k is outside T_k, which only loosens the Gershgorin enclosure.
            const int kprev = k > 16 ? k - 16 : k;
#ifdef FKMC_KPM_TIMING
            t_tmp = clock64();
#endif
            for (int task = warp; task < 4; task += nwarps) {
                const int kk = (task < 2) ? k : kprev;
                const double v = warp_ritz(ab, kk, (task & 1) ? kk - 1 : 0, gl, gh, lane);
                if (lane == 0) msc[task] = v;
            }
            __syncthreads();
            e_min = msc[0];
            e_max = msc[1];
            const double tol = 8.0 * DBL_EPSILON * hscale;
            if (k > 16 && fabs(e_min - msc[2]) <= tol && fabs(e_max - msc[3]) <= tol) converged = true;
            __syncthreads();
#ifdef FKMC_KPM_TIMING
            t_ritz += clock64() - t_tmp;
#endif
            if (converged || last) break;
        }
    }
#ifdef FKMC_KPM_TIMING
    const long long t_lanczos_end = clock64();
#endif
    if (!converged && k == kcap && k < N && tid == 0) atomicOr(P.flag, 2);
    if (P.steps && tid == 0) P.steps[b] = k;

    // =========================== 2. moments ===========================
    const double a = (e_max - e_min) / 2., bsh = (e_max + e_min) / 2.;
    double part = 0.0;
    for (int i = tid; i < N; i += T) {
        const double x = (xd[i] - bsh) / a;
        xd[i] = x;
        part += x;
    }
    const double trx = block_sum(part, red);  // (two barriers: also orders the Lanczos reads before reuse)
    double sv[FKMC_MAX_Z];
#pragma unroll
    for (int z = 0; z < FKMC_MAX_Z; ++z) sv[z] = z < Z ? P.slot_val[z] / a : 0.0;

    double tr[HALF + 1], d01[HALF + 1], d11[HALF + 1];
#pragma unroll
    for (int m = 0; m <= HALF; ++m) tr[m] = d01[m] = d11[m] = 0.0;
    if constexpr (KIND != 0) {
        // Local-patch recursion.  The patch (PW x PW, flattened with row stride PW) is split into 32 contiguous runs of R
        // elements, one per lane (R odd -> conflict-free shared-memory access); lane 31 also owns the last element when
        // 32 R = PP - 1.  A lane keeps its runs of T_{m-1} e_j and T_{m-2} e_j in registers, so the +-1 neighbours come from
        // registers (run ends: one shuffle) and only the +-PW (+-(PW+1)) neighbours are read from shared memory.
        constexpr int H = HALF, PW = 2 * H + 1, PP = PW * PW, GUARD = PW + 1, PVp = (PP + 2 * GUARD + 1) & ~1;
        constexpr int R0 = (PP - 1 + 31) / 32, R = (R0 % 2) ? R0 : R0 + 1;   // run length (odd)
        constexpr bool TAIL = (32 * R == PP - 1);                             // one extra element for lane 31
        constexpr int RE = R + (TAIL ? 1 : 0);
        constexpr int KC = H * PW + H, LANE_C = KC / R, E_C = KC % R;         // the centre of the patch
        const int L = P.L;
        double* const sb0 = vec + (size_t)warp * 2 * PVp + GUARD;  // shared copies, index k in [0, PP), zero guard zones around
        double* const sb1 = sb0 + PVp;
        const double svt = sv[0];                                   // nearest-neighbour hopping / a
        const double svp = (KIND == FKMC_TRIANGULAR) ? sv[4] : 0.0;  // (x-1,y-1)/(x+1,y+1) hopping / a
        for (int i = lane; i < 2 * PVp; i += 32) sb0[i - GUARD] = 0.0;  // both buffers incl. guards (contiguous), once
        __syncwarp();
        const int kbase = R * lane;
        for (int j = warp; j < N; j += nwarps) {
            const int y0 = j / L, x0 = j - y0 * L;
            const bool jeven = ((y0 + x0) & 1) == 0;  // honeycomb: sublattice A hops up (y+1), B hops down
            double xdp[RE], va[RE], vb[RE];           // diagonal of X on the patch; T_{m-2} e_j (va) and T_{m-1} e_j (vb)
            bool up[RE];                              // honeycomb: this site's vertical bond goes to +PW
#pragma unroll
            for (int e = 0; e < RE; ++e) {
                const int kq = kbase + e;
                const bool live = (e < R) ? (kq < PP) : (TAIL && lane == 31);
                const int dy = kq / PW - H, dx = kq % PW - H;
                int yy = y0 + dy, xx = x0 + dx;
                yy += (yy < 0) ? L : 0; yy -= (yy >= L) ? L : 0;
                xx += (xx < 0) ? L : 0; xx -= (xx >= L) ? L : 0;
                xdp[e] = live ? xd[yy * L + xx] : 0.0;
                up[e] = (((dy + dx) & 1) == 0) == jeven;
                // v0 = e_j, v1 = X e_j (column j of the symmetric X)
                const int off = kq - KC;
                va[e] = (live && off == 0) ? 1.0 : 0.0;
                double v1 = 0.0;
                if (live) {
                    if (off == 0) v1 = xd[j];
                    else if (off == 1 || off == -1) v1 = svt;
                    else if (KIND == FKMC_HONEYCOMB) { if (off == (jeven ? PW : -PW)) v1 = svt; }
                    else if (off == PW || off == -PW) v1 = svt;
                    else if (KIND == FKMC_TRIANGULAR && (off == PW + 1 || off == -PW - 1)) v1 = svp;
                }
                vb[e] = v1;
                if (live) sb1[kq] = v1;
            }
            __syncwarp();
            // one recursion step: vold <- 2 X vcur - vold (in registers), published to the shared buffer snew
            auto step = [&](double (&vold)[RE], const double (&vcur)[RE], const double* __restrict__ scur, double* __restrict__ snew,
                            bool need_dots, double& s01, double& s11, double& centre) {
                // run-end neighbours from the adjacent lanes
                double left = __shfl_up_sync(0xffffffffu, vcur[R - 1], 1);
                double right = __shfl_down_sync(0xffffffffu, vcur[0], 1);
                if (lane == 0) left = 0.0;
                if (lane == 31) right = TAIL ? vcur[RE - 1] : 0.0;
#pragma unroll
                for (int e = 0; e < RE; ++e) {
                    const int kq = kbase + e;
                    const bool live = (e < R) ? (kq < PP) : (TAIL && lane == 31);
                    if (live) {
                        const double c1 = vcur[e];
                        const double lft = (e == 0) ? left : vcur[e - 1];
                        const double rgt = (e == R - 1) ? right : ((e < R - 1) ? vcur[e + 1] : 0.0);
                        double nb = lft + rgt;
                        if (KIND == FKMC_HONEYCOMB) nb += scur[up[e] ? kq + PW : kq - PW];
                        else nb += scur[kq - PW] + scur[kq + PW];
                        double sacc = fma(xdp[e], c1, svt * nb);
                        if (KIND == FKMC_TRIANGULAR) sacc = fma(svp, scur[kq - PW - 1] + scur[kq + PW + 1], sacc);
                        const double vn = 2. * sacc - vold[e];
                        vold[e] = vn;
                        snew[kq] = vn;
                        if (need_dots) {
                            s01 = fma(c1, vn, s01);
                            s11 = fma(vn, vn, s11);
                        }
                    }
                }
                centre = (lane == LANE_C) ? vold[E_C] : 0.0;
                __syncwarp();
            };
#pragma unroll
            for (int m = 2; m <= HALF; ++m) {
                double s01 = 0.0, s11 = 0.0, centre = 0.0;
                const bool need_dots = (2 * m - 1 >= HALF);
                if ((m & 1) == 0) step(va, vb, sb1, sb0, need_dots, s01, s11, centre);   // va <- T_m e_j
                else step(vb, va, sb0, sb1, need_dots, s01, s11, centre);                 // vb <- T_m e_j
                tr[m] += centre;
                d01[m] += s01;
                d11[m] += s11;
            }
        }
    } else
    if (warp < P.nwarps_cols) {
        double* v0 = vec + (size_t)warp * 2 * Nv;
        double* v1 = v0 + Nv;
        for (int j = warp; j < N; j += P.nwarps_cols) {
            // v0 = e_j, v1 = X e_j (column j of the symmetric X)
            for (int i = lane; i < Nv; i += 32) { v0[i] = 0.0; v1[i] = 0.0; }
            __syncwarp();
            if (lane == 0) {
                v0[j] = 1.0;
                v1[j] = xd[j];
                for (int z = 0; z < Z; ++z) {
                    const int nb = nidx[z * N + j];
                    if (nb < N) v1[nb] += sv[z];
                }
            }
            __syncwarp();
#pragma unroll
            for (int m = 2; m <= HALF; ++m) {
                // v0 <- 2 X v1 - v0 (in place), then swap roles
                double s01 = 0.0, s11 = 0.0;
                const bool need_dots = (2 * m - 1 >= HALF);
                for (int i = lane; i < N; i += 32) {
                    double s = xd[i] * v1[i];
#pragma unroll
                    for (int z = 0; z < FKMC_MAX_Z; ++z)
                        if (z < Z) s = fma(sv[z], v1[nidx[z * N + i]], s);
                    const double vn = 2. * s - v0[i];
                    v0[i] = vn;
                    if (need_dots) {
                        s01 = fma(v1[i], vn, s01);
                        s11 = fma(vn, vn, s11);
                    }
                }
                __syncwarp();
                if (lane == (j & 31)) tr[m] += v0[j];
                d01[m] += s01;
                d11[m] += s11;
                double* t = v0; v0 = v1; v1 = t;
            }
            __syncwarp();
        }
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
#ifdef FKMC_KPM_TIMING
    if (tid == 0 && b == 0)
        printf("kpm timing (cycles): lanczos total %lld (ritz %lld, steps %d) moments %lld\n", t_lanczos_end - t_start, t_ritz, k,
               (long long)clock64() - t_lanczos_end);
#endif
    // =========================== 3. coefficients and logZ ===========================
    double* mom = msc + 8;  // [M] (M <= 32)
    if (tid == 0) {
        bool is_set[2 * FKMC_MAX_HALF];
        for (int m = 0; m < M; ++m) { is_set[m] = false; mom[m] = 0.0; }
        mom[0] = 1.0; is_set[0] = true;
        mom[1] = trx / N; is_set[1] = true;
        for (int m = 2; m <= HALF; ++m) {
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            for (int w = 0; w < nwarps; ++w) {
                t0 += acc[(w * 3 + 0) * (HALF + 1) + m];
                t1 += acc[(w * 3 + 1) * (HALF + 1) + m];
                t2 += acc[(w * 3 + 2) * (HALF + 1) + m];
            }
            if (!is_set[m]) { mom[m] = t0 / N; is_set[m] = true; }
            int kk = 2 * m - 1;
            if (kk < M && kk >= HALF) {
                mom[kk] = (t1 * 2. - trx) / N; is_set[kk] = true;
                if (kk != M - 1) { ++kk; mom[kk] = (t2 / N * 2. - 1.0); is_set[kk] = true; }
            }
        }
    }
    for (int i = tid; i < G; i += T) Fg[i] = N * log(1. + exp(-P.beta * (a * P.lobatto[i] + bsh)));
    __syncthreads();
    if (tid < M) {
        const double* Tm = P.chebt + (size_t)tid * G;
        double s = 0.0;
        for (int i = 0; i < G - 1; ++i) s += (Fg[i + 1] * Tm[i + 1] + Fg[i] * Tm[i]) * P.dtheta[i];
        acc[tid] = s * 0.5;  // acc is free now: reuse for the coefficients c_m
    }
    __syncthreads();
    if (tid == 0) {
        double s = acc[0];
        for (int m = 1; m < M; ++m) s += 2. * acc[m] * mom[m];
        P.logz[b] = s;
        P.ab[(size_t)b * 4 + 0] = e_min;
        P.ab[(size_t)b * 4 + 1] = e_max;
        P.ab[(size_t)b * 4 + 2] = a;
        P.ab[(size_t)b * 4 + 3] = bsh;
    }
    if (tid < M && P.moments) P.moments[(size_t)b * M + tid] = mom[tid];
}

}  // namespace

// Chebyshev tables (include/fk_mc/chebyshev.hpp:21-34): theta_i uniform on [0,1], x_i = -cos(pi theta_i),
// T_k(x_i) = cos(k acos x_i).  Cached per (M, G).
int fkmc_prepare_cheb(fkmc_ctx* ctx, int M, int G) {
    if (ctx->cheb_M == M && ctx->cheb_G == G && ctx->d_chebt) return FKMC_OK;
    if (M < 2 || M % 2 || M > 2 * FKMC_MAX_HALF || G < 2 || G > 4096)
        return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: need even 2 <= M <= 32 and 2 <= G <= 4096");
    std::vector<double> theta(G), x(G), T((size_t)M * G), dth(G - 1);
    for (int i = 0; i < G; ++i) {
        theta[i] = (i == G - 1) ? 1.0 : double(i) * (1.0 / double(G - 1));
        x[i] = -std::cos(M_PI * theta[i]);
        for (int k = 0; k < M; ++k) T[(size_t)k * G + i] = std::cos(k * std::acos(x[i]));
    }
    for (int i = 0; i < G - 1; ++i) dth[i] = theta[i + 1] - theta[i];
    if (ctx->d_chebt) { cudaFree(ctx->d_chebt); cudaFree(ctx->d_lobatto); cudaFree(ctx->d_dtheta); ctx->d_chebt = nullptr; }
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_chebt, sizeof(double) * M * G));
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_lobatto, sizeof(double) * G));
    FKMC_CUDA(ctx, cudaMalloc(&ctx->d_dtheta, sizeof(double) * (G - 1)));
    FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_chebt, T.data(), sizeof(double) * M * G, cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_lobatto, x.data(), sizeof(double) * G, cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_dtheta, dth.data(), sizeof(double) * (G - 1), cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->cheb_M = M;
    ctx->cheb_G = G;
    return FKMC_OK;
}

template <int HALF, int KIND>
static int launch_kpm_patch(fkmc_ctx* ctx, kpm_args& P, int B) {
    constexpr int PW = 2 * HALF + 1, PP = PW * PW, GUARD = PW + 1, PVp = (PP + 2 * GUARD + 1) & ~1;
    const int N = P.N, Nv = N + 1, nw = 8;
    const size_t lanczos_need = 3 * (size_t)Nv + 2 * KPM_KMAX + 6;
    size_t vec_doubles = std::max((size_t)nw * 2 * PVp, lanczos_need);
    vec_doubles += vec_doubles & 1;
    const size_t fixed = sizeof(double) * ((size_t)N + 48 + 64 + ((P.G + 1) & ~1) + (size_t)nw * 3 * (HALF + 1));
    const size_t smem = fixed + sizeof(double) * vec_doubles + sizeof(unsigned short) * (size_t)P.Z * N + 16;
    if (smem > ctx->smem_optin - 1024) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: lattice too large for the shared-memory kernel");
    if (N > 8 * nw * 32) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: lattice too large for the Lanczos register tiling");
    P.nwarps_cols = nw;
    P.vec_doubles = (int)vec_doubles;
    FKMC_CUDA(ctx, cudaFuncSetAttribute(kpm_kernel<HALF, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kpm_kernel<HALF, KIND><<<B, nw * 32, smem, ctx->stream>>>(P);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

template <int HALF>
static int launch_kpm_t(fkmc_ctx* ctx, kpm_args& P, int B) {
    // local-patch variant when the lattice is 2-D and the m <= HALF hop neighbourhood does not wrap onto itself
    if (HALF >= 2 && HALF <= 10 && P.L >= 2 * HALF + 1 && !ctx->kpm_force_generic) {
        if constexpr (HALF >= 2 && HALF <= 10) {
            if (ctx->kind == FKMC_CUBIC2D) return launch_kpm_patch<HALF, FKMC_CUBIC2D>(ctx, P, B);
            if (ctx->kind == FKMC_TRIANGULAR) return launch_kpm_patch<HALF, FKMC_TRIANGULAR>(ctx, P, B);
            if (ctx->kind == FKMC_HONEYCOMB) return launch_kpm_patch<HALF, FKMC_HONEYCOMB>(ctx, P, B);
        }
    }
    const int N = P.N, Nv = N + 1;
    // shared memory: fixed part + 2 vectors per column-warp; need >= 2 column warps (Lanczos uses 4 buffers)
    const size_t budget = ctx->smem_optin - 1024;
    int nw = 12;
    size_t smem = 0, vec_doubles = 0;
    const size_t lanczos_need = 3 * (size_t)Nv + 2 * KPM_KMAX + 6;
    for (; nw >= 2; --nw) {
        const size_t fixed = sizeof(double) * ((size_t)N + 48 + 64 + ((P.G + 1) & ~1) + (size_t)nw * 3 * (HALF + 1));
        vec_doubles = std::max((size_t)nw * 2 * Nv, lanczos_need);
        vec_doubles += vec_doubles & 1;
        const size_t idx = sizeof(unsigned short) * (size_t)P.Z * N + 16;
        smem = fixed + sizeof(double) * vec_doubles + idx;
        if (smem <= budget) break;
    }
    if (nw < 2) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: lattice too large for the shared-memory kernel");
    if (N > 8 * nw * 32) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: lattice too large for the Lanczos register tiling");
    P.nwarps_cols = nw;
    P.vec_doubles = (int)vec_doubles;
    FKMC_CUDA(ctx, cudaFuncSetAttribute(kpm_kernel<HALF, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kpm_kernel<HALF, 0><<<B, nw * 32, smem, ctx->stream>>>(P);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

int fkmc_launch_kpm(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, int M, int G,
                    double* d_moments, double* d_ab, double* d_logz) {
    int rc = fkmc_prepare_cheb(ctx, M, G);
    if (rc) return rc;
    if (ctx->N > 65535) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: N > 65535");
    fkmc_prof_scope ps(ctx, "kpm");
    kpm_args P{};
    P.f = d_f; P.nbr_idx = ctx->d_nbr_idx; P.N = ctx->N; P.Z = ctx->Z; P.M = M; P.G = G; P.L = ctx->L;
    P.U = U; P.mu_c = mu_c; P.beta = beta;
    // per-slot hopping constants (all supported stencils are uniform per slot; checked here)
    for (int z = 0; z < FKMC_MAX_Z; ++z) P.slot_val[z] = 0.0;
    for (int z = 0; z < ctx->Z; ++z) {
        bool have = false;
        for (int i = 0; i < ctx->N; ++i) {
            if (ctx->h_nbr_idx[(size_t)z * ctx->N + i] >= ctx->N) continue;
            const double v = ctx->h_nbr_val[(size_t)z * ctx->N + i];
            if (!have) { P.slot_val[z] = v; have = true; }
            else if (v != P.slot_val[z]) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: non-uniform hopping within a stencil slot");
        }
    }
    P.chebt = ctx->d_chebt; P.lobatto = ctx->d_lobatto; P.dtheta = ctx->d_dtheta;
    P.moments = d_moments; P.ab = d_ab; P.logz = d_logz; P.flag = ctx->d_flag; P.steps = ctx->d_kpm_steps;
    P.kmax = ctx->lanczos_cap > 0 ? std::min(ctx->lanczos_cap, KPM_KMAX) : KPM_KMAX;
    ctx->kpm_state_written = false;
    if (fkmc_kpm2d_applicable(ctx, M)) return fkmc_launch_kpm2d(ctx, d_f, B, U, mu_c, beta, M, G, P.slot_val, d_moments, d_ab, d_logz);
    switch (M / 2) {
#define FKMC_KPM_CASE(H) case H: return launch_kpm_t<H>(ctx, P, B);
        FKMC_KPM_CASE(1) FKMC_KPM_CASE(2) FKMC_KPM_CASE(3) FKMC_KPM_CASE(4) FKMC_KPM_CASE(5) FKMC_KPM_CASE(6)
        FKMC_KPM_CASE(7) FKMC_KPM_CASE(8) FKMC_KPM_CASE(9) FKMC_KPM_CASE(10) FKMC_KPM_CASE(11) FKMC_KPM_CASE(12)
        FKMC_KPM_CASE(13) FKMC_KPM_CASE(14) FKMC_KPM_CASE(15) FKMC_KPM_CASE(16)
#undef FKMC_KPM_CASE
    }
    return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: unsupported M");
}
