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
    int L;                  // linear lattice size (patch variant)
    int nwarps_cols;        // warps that run the column recursion
    int vec_doubles;        // size of the shared vector region (>= 2 Nv per column warp and >= the Lanczos need)
};

// number of eigenvalues of the k x k Lanczos tridiagonal (al[0..k-1], off-diagonals be[1..k-1]) below x
__device__ __forceinline__ int lanczos_sturm(const double* al, const double* be, int k, double x) {
    double pm1 = 1.0, p = al[0] - x;
    if (p == 0.0) p = -DBL_EPSILON;
    bool neg = p < 0.0;
    int cnt = neg ? 1 : 0;
    for (int i = 1; i < k; ++i) {
        const double bb = be[i];
        double pn = fma(al[i] - x, p, -(bb * bb * pm1));
        if (pn == 0.0) pn = -DBL_EPSILON * p;
        const bool nneg = pn < 0.0;
        cnt += (nneg != neg) ? 1 : 0;
        neg = nneg;
        pm1 = p;
        p = pn;
        if ((i & 7) == 0) {
            const double m = fmax(fabs(p), fabs(pm1));
            if (m > 1.157920892373162e77) { p *= 8.636168555094445e-78; pm1 *= 8.636168555094445e-78; }
            else if (m < 8.636168555094445e-78) { p *= 1.157920892373162e77; pm1 *= 1.157920892373162e77; }
        }
    }
    return cnt;
}

// idx-th eigenvalue of the Lanczos tridiagonal by 32-way multisection (whole warp participates)
__device__ __forceinline__ double warp_ritz(const double* al, const double* be, int k, int idx, double lo, double hi, int lane) {
    const double pad = 8.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) + DBL_MIN;
    double a = lo - pad, c = hi + pad;
    for (int round = 0; round < 14; ++round) {
        const double h = (c - a) * (1.0 / 33.0);
        if (!(h > 2.0 * DBL_EPSILON * fmax(fabs(a), fabs(c)) * (1.0 / 33.0))) break;
        const double x = a + h * (double)(lane + 1);
        const bool above = lanczos_sturm(al, be, k, x) > idx;
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
    // vectors live in the first three column buffers; alpha/beta in the fourth
    double* lv = vec;            // v_k
    double* lp = vec + Nv;       // v_{k-1}
    double* lw = vec + 2 * Nv;   // w
    double* al = vec + 3 * Nv;   // [KPM_KMAX]
    double* be = al + KPM_KMAX;  // [KPM_KMAX + 1]
    {
        double part = 0.0;
        for (int i = tid; i < Nv; i += T) {
            double v = 0.0;
            if (i < N) {
                unsigned h = (unsigned)i * 2654435761u + 0x9e3779b9u;
                h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
                v = (double)h * (1.0 / 4294967296.0) - 0.5;
            }
            lv[i] = v; lp[i] = 0.0; lw[i] = 0.0;
            part = fma(v, v, part);
        }
        const double nrm = sqrt(block_sum(part, red));
        for (int i = tid; i < N; i += T) lv[i] /= nrm;
        if (tid == 0) be[0] = 0.0;
        __syncthreads();
    }
    const int kcap = min(KPM_KMAX, N);
    double e_min = 0.0, e_max = 0.0, prev_min = 0.0, prev_max = 0.0, gl = DBL_MAX, gh = -DBL_MAX, hscale = 0.0;
    bool have_prev = false, converged = false;
    int k = 0;
    double beta_k = 0.0;
    while (k < kcap) {
        double part = 0.0;
        for (int i = tid; i < N; i += T) {
            double s = xd[i] * lv[i];
            for (int z = 0; z < Z; ++z) s = fma(P.slot_val[z], lv[nidx[z * N + i]], s);
            s = fma(-beta_k, lp[i], s);
            lw[i] = s;
            part = fma(s, lv[i], part);
        }
        const double alpha = block_sum(part, red);
        part = 0.0;
        for (int i = tid; i < N; i += T) {
            const double s = fma(-alpha, lv[i], lw[i]);
            lw[i] = s;
            part = fma(s, s, part);
        }
        const double nb = sqrt(block_sum(part, red));
        if (tid == 0) { al[k] = alpha; be[k + 1] = nb; }
        // Gershgorin enclosure of the Lanczos tridiagonal (all threads keep it in registers)
        gl = fmin(gl, alpha - beta_k - nb);
        gh = fmax(gh, alpha + beta_k + nb);
        hscale = fmax(hscale, fabs(alpha) + nb);
        ++k;
        const bool breakdown = nb <= 1e-13 * hscale;
        if (!breakdown) {
            const double inv = 1.0 / nb;
            for (int i = tid; i < N; i += T) {
                const double s = lw[i] * inv;
                lw[i] = lp[i];  // recycled as scratch next step
                lp[i] = lv[i];
                lv[i] = s;
            }
        }
        beta_k = nb;
        __syncthreads();
        const bool last = breakdown || k == kcap;
        if (last || (k >= 32 && (k & 15) == 0)) {
            // the enclosure must not count the (dropped) last off-diagonal: be[k] is outside T_k
            if (warp == 0) {
                const double v = warp_ritz(al, be, k, 0, gl, gh, lane);
                if (lane == 0) msc[0] = v;
            } else if (warp == 1) {
                const double v = warp_ritz(al, be, k, k - 1, gl, gh, lane);
                if (lane == 0) msc[1] = v;
            }
            __syncthreads();
            e_min = msc[0];
            e_max = msc[1];
            const double tol = 8.0 * DBL_EPSILON * hscale;
            if (have_prev && fabs(e_min - prev_min) <= tol && fabs(e_max - prev_max) <= tol) converged = true;
            prev_min = e_min; prev_max = e_max; have_prev = true;
            __syncthreads();
            if (converged || last) break;
        }
    }
    if (!converged && k == kcap && k < N && tid == 0) atomicOr(P.flag, 2);

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
        constexpr int H = HALF, PW = 2 * H + 1, PP = PW * PW, GUARD = PW + 1, PVp = (PP + 2 * GUARD + 1) & ~1, NE = (PP + 31) / 32;
        constexpr int CEN = GUARD + H * PW + H;
        const int L = P.L;
        double* const vbase = vec + (size_t)warp * 2 * PVp;
        const double svt = sv[0];                                   // nearest-neighbour hopping / a
        const double svp = (KIND == FKMC_TRIANGULAR) ? sv[4] : 0.0;  // (x-1,y-1)/(x+1,y+1) hopping / a
        int dyl[NE], dxl[NE];
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            const int kq = lane + 32 * i;
            dyl[i] = kq / PW - H;
            dxl[i] = kq % PW - H;
        }
        for (int j = warp; j < N; j += nwarps) {
            const int y0 = j / L, x0 = j - y0 * L;
            double* v0 = vbase;  // (the roles are swapped HALF-1 times per column: always restart from the base layout)
            double* v1 = vbase + PVp;
            double xdp[NE];
#pragma unroll
            for (int i = 0; i < NE; ++i) {
                int yy = y0 + dyl[i], xx = x0 + dxl[i];
                yy += (yy < 0) ? L : 0; yy -= (yy >= L) ? L : 0;
                xx += (xx < 0) ? L : 0; xx -= (xx >= L) ? L : 0;
                xdp[i] = (lane + 32 * i < PP) ? xd[yy * L + xx] : 0.0;
            }
            for (int i = lane; i < 2 * PVp; i += 32) v0[i] = 0.0;  // both buffers are contiguous
            __syncwarp();
            const bool jeven = ((y0 + x0) & 1) == 0;  // honeycomb: sublattice A hops up (y+1), B hops down
            if (lane == 0) {
                v0[CEN] = 1.0;
                v1[CEN] = xd[j];
                v1[CEN - 1] = svt;
                v1[CEN + 1] = svt;
                if (KIND == FKMC_HONEYCOMB) {
                    v1[jeven ? CEN + PW : CEN - PW] = svt;
                } else {
                    v1[CEN - PW] = svt;
                    v1[CEN + PW] = svt;
                    if (KIND == FKMC_TRIANGULAR) { v1[CEN - PW - 1] = svp; v1[CEN + PW + 1] = svp; }
                }
            }
            __syncwarp();
#pragma unroll
            for (int m = 2; m <= HALF; ++m) {
                double s01 = 0.0, s11 = 0.0;
                const bool need_dots = (2 * m - 1 >= HALF);
#pragma unroll
                for (int i = 0; i < NE; ++i) {
                    const int kq = lane + 32 * i;
                    if (kq < PP) {
                        const int kk = kq + GUARD;
                        const double c1 = v1[kk];
                        double nb = v1[kk - 1] + v1[kk + 1];
                        if (KIND == FKMC_HONEYCOMB) {
                            const bool even = (((dyl[i] + dxl[i]) & 1) == 0) == jeven;
                            nb += v1[even ? kk + PW : kk - PW];
                        } else {
                            nb += v1[kk - PW] + v1[kk + PW];
                        }
                        double sacc = fma(xdp[i], c1, svt * nb);
                        if (KIND == FKMC_TRIANGULAR) sacc = fma(svp, v1[kk - PW - 1] + v1[kk + PW + 1], sacc);
                        const double vn = 2. * sacc - v0[kk];
                        v0[kk] = vn;
                        if (need_dots) {
                            s01 = fma(c1, vn, s01);
                            s11 = fma(vn, vn, s11);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) tr[m] += v0[CEN];
                d01[m] += s01;
                d11[m] += s11;
                double* tswap = v0; v0 = v1; v1 = tswap;
            }
            __syncwarp();
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
    const size_t lanczos_need = 3 * (size_t)Nv + 2 * KPM_KMAX + 2;
    size_t vec_doubles = std::max((size_t)nw * 2 * PVp, lanczos_need);
    vec_doubles += vec_doubles & 1;
    const size_t fixed = sizeof(double) * ((size_t)N + 48 + 64 + ((P.G + 1) & ~1) + (size_t)nw * 3 * (HALF + 1));
    const size_t smem = fixed + sizeof(double) * vec_doubles + sizeof(unsigned short) * (size_t)P.Z * N + 16;
    if (smem > ctx->smem_optin - 1024) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: lattice too large for the shared-memory kernel");
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
    const size_t lanczos_need = 3 * (size_t)Nv + 2 * KPM_KMAX + 2;
    for (; nw >= 2; --nw) {
        const size_t fixed = sizeof(double) * ((size_t)N + 48 + 64 + ((P.G + 1) & ~1) + (size_t)nw * 3 * (HALF + 1));
        vec_doubles = std::max((size_t)nw * 2 * Nv, lanczos_need);
        vec_doubles += vec_doubles & 1;
        const size_t idx = sizeof(unsigned short) * (size_t)P.Z * N + 16;
        smem = fixed + sizeof(double) * vec_doubles + idx;
        if (smem <= budget) break;
    }
    if (nw < 2) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: lattice too large for the shared-memory kernel");
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
    P.moments = d_moments; P.ab = d_ab; P.logz = d_logz; P.flag = ctx->d_flag;
    switch (M / 2) {
#define FKMC_KPM_CASE(H) case H: return launch_kpm_t<H>(ctx, P, B);
        FKMC_KPM_CASE(1) FKMC_KPM_CASE(2) FKMC_KPM_CASE(3) FKMC_KPM_CASE(4) FKMC_KPM_CASE(5) FKMC_KPM_CASE(6)
        FKMC_KPM_CASE(7) FKMC_KPM_CASE(8) FKMC_KPM_CASE(9) FKMC_KPM_CASE(10) FKMC_KPM_CASE(11) FKMC_KPM_CASE(12)
        FKMC_KPM_CASE(13) FKMC_KPM_CASE(14) FKMC_KPM_CASE(15) FKMC_KPM_CASE(16)
#undef FKMC_KPM_CASE
    }
    return fkmc_set_error(ctx, FKMC_ERR_INVALID, "KPM: unsupported M");
}
