// Eigenvalues of symmetric tridiagonals by Sturm bisection (one CTA per matrix, one thread per
// eigenvalue) with the free-energy reduction fused in.
//
// Replaces the implicit-QR stage of Eigen::SelfAdjointEigenSolver (call site
// src/configuration.cpp:213) and the cache fill of configuration_t::calc_ed
// (src/configuration.cpp:226-244): cached_exp = e^{beta eps}, cached_fermi = 1/(1+e^{beta eps}),
// logZ = sum_i [log(e^{beta eps_0} + e^{-beta (eps_i - eps_0)}) - beta eps_0].
// Also measure_energy::accumulate (src/measures/energy.cpp:6-26) from a cached spectrum.
//
// Sturm count in product form p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (no division; one
// dependent FMA per row) with periodic power-of-two rescaling; (d_i, e_{i-1}^2) pairs are
// broadcast from shared memory.
#include <algorithm>
#include <cfloat>

#include "common.cuh"

#ifndef FKMC_EIG_PREDSTOP
#define FKMC_EIG_PREDSTOP 1  // stop the Newton phase on the predicted (quadratic) error instead of a confirming evaluation
#endif
#ifndef FKMC_EIG_BISECT
#define FKMC_EIG_BISECT 6  // bisection steps every lane takes before the Newton phase (measured per 1024 matrices of N = 1024 with the grouped recurrences: 5 -> 4.88 ms, 6 -> 4.74, 8 -> 5.06, 10 -> 5.23)
#endif

namespace {

// number of eigenvalues < x.  de[i] = (d_i, e_{i-1}^2), e_{-1} = 0.  Guarded form: an exact zero is nudged off zero, so the
// recurrence survives e_i = 0 (a decoupled block) right after a zero.  Slow path of sturm_count below.
__device__ __noinline__ int sturm_count_guarded(const double2* __restrict__ de, int n, double x) {
    double pm1 = 1.0, p = de[0].x - x;
    if (p == 0.0) p = -DBL_EPSILON;
    bool neg = p < 0.0;
    int cnt = neg ? 1 : 0;
    for (int i0 = 1; i0 < n; i0 += 8) {
        const int i1 = min(i0 + 8, n);
        for (int i = i0; i < i1; ++i) {
            const double2 q = de[i];
            double pn = fma(q.x - x, p, -(q.y * pm1));
            if (pn == 0.0) pn = -DBL_EPSILON * p;
            const bool nneg = pn < 0.0;
            cnt += (nneg != neg) ? 1 : 0;
            neg = nneg;
            pm1 = p;
            p = pn;
        }
        // rescale by a power of two when the pair drifts out of [2^-256, 2^256]
        const double m = fmax(fabs(p), fabs(pm1));
        if (m > 1.157920892373162e77) { p *= 8.636168555094445e-78; pm1 *= 8.636168555094445e-78; }
        else if (m < 8.636168555094445e-78) { p *= 1.157920892373162e77; pm1 *= 1.157920892373162e77; }
    }
    return cnt;
}

// Fast form: three FP64 instructions per row (subtract, multiply, FMA); the sign of every p_i is shifted into a history word by one
// funnel shift on the integer pipe and the sign changes of a group of 16 rows are counted at once (popc of the word against itself
// shifted by one) - the kernel is issue-bound, so instructions per row are what matters.  The pair (p, p_{-1}) grows or shrinks by at
// most 5x per row (|d_i - x| <= 4, e^2 <= 1 after scaling), so a range check per group keeps it away from overflow; shrinking is only
// bounded by the e_i^2, see STURM_TINY below.
// An isolated exact zero needs no care (p_{i+1} = -e_i^2 p_{i-1} then has the sign opposite to p_{i-1}: one sign change whichever
// sign the zero is given); only a zero followed by e_i = 0 sticks, which shows as p = p_{-1} = 0 at the next rescaling check and is
// sent to the guarded form.
#ifndef FKMC_STURM_GROUP
#define FKMC_STURM_GROUP 24  // measured at N = 1024, 1024 matrices: 8 -> 5.75 ms, 16 -> 4.74, 24 -> 4.53, 31 -> 4.59
#endif
// A pair that ends a group below 2^-700 may have passed through the denormal range inside it (it grows by at most 5x per row, so a
// pair that was below 2^-1022 anywhere in a group of <= 31 rows ends it below 2^-950) and is re-evaluated by the guarded form; a pair
// whose maximum stays normal loses nothing: a denormal member then is negligible against the other term of the recurrence.
constexpr double STURM_TINY = 0x1p-700;
constexpr int STURM_GROUP = FKMC_STURM_GROUP;  // <= 31: the history word holds the group and the sign before it

__device__ __forceinline__ int sturm_count(const double2* __restrict__ de, int n, double x) {
    double pm1 = 1.0, p = de[0].x - x;
    unsigned hist = (unsigned)__double2hiint(p) >> 31;  // bit k = sign of p_{i-k}
    int cnt = (int)hist;
    int i = 1;
    for (; i + STURM_GROUP <= n; i += STURM_GROUP) {
        hist &= 1u;
#pragma unroll
        for (int k = 0; k < STURM_GROUP; ++k) {
            const double2 q = de[i + k];
            const double pn = fma(q.x - x, p, -(q.y * pm1));
            hist = __funnelshift_l((unsigned)__double2hiint(pn), hist, 1);
            pm1 = p;
            p = pn;
        }
        cnt += __popc((hist ^ (hist >> 1)) & ((1u << STURM_GROUP) - 1u));
        const double m = fmax(fabs(p), fabs(pm1));
        if (m > 1.157920892373162e77) { p *= 8.636168555094445e-78; pm1 *= 8.636168555094445e-78; }
        else if (m < 8.636168555094445e-78) {
            if (m < STURM_TINY) return sturm_count_guarded(de, n, x);
            p *= 1.157920892373162e77; pm1 *= 1.157920892373162e77;
        }
    }
    if (i < n) {  // last, shorter group
        const int r = n - i;
        hist &= 1u;
        for (; i < n; ++i) {
            const double2 q = de[i];
            const double pn = fma(q.x - x, p, -(q.y * pm1));
            hist = __funnelshift_l((unsigned)__double2hiint(pn), hist, 1);
            pm1 = p;
            p = pn;
        }
        cnt += __popc((hist ^ (hist >> 1)) & ((1u << r) - 1u));
        if (fmax(fabs(p), fabs(pm1)) < STURM_TINY) return sturm_count_guarded(de, n, x);
    }
    return cnt;
}

// Sturm count plus the characteristic polynomial p_n(x) and its derivative (same recurrence, differentiated:
// p'_i = (d_i - x) p'_{i-1} - p_{i-1} - e_{i-1}^2 p'_{i-2}), jointly rescaled, for a safeguarded Newton step.  Guarded form.
__device__ __noinline__ int sturm_newton_guarded(const double2* __restrict__ de, int n, double x, double& pn_out, double& dpn_out) {
    double pm1 = 1.0, p = de[0].x - x, dpm1 = 0.0, dp = -1.0;
    if (p == 0.0) p = -DBL_EPSILON;
    bool neg = p < 0.0;
    int cnt = neg ? 1 : 0;
    for (int i0 = 1; i0 < n; i0 += 8) {
        const int i1 = min(i0 + 8, n);
        for (int i = i0; i < i1; ++i) {
            const double2 q = de[i];
            const double t = q.x - x;
            double pn = fma(t, p, -(q.y * pm1));
            const double dpn = fma(t, dp, -fma(q.y, dpm1, p));
            if (pn == 0.0) pn = -DBL_EPSILON * p;
            const bool nneg = pn < 0.0;
            cnt += (nneg != neg) ? 1 : 0;
            neg = nneg;
            pm1 = p; p = pn;
            dpm1 = dp; dp = dpn;
        }
        const double m = fmax(fmax(fabs(p), fabs(pm1)), fmax(fabs(dp), fabs(dpm1)) * 0x1p-60);
        if (m > 1.157920892373162e77) {
            p *= 8.636168555094445e-78; pm1 *= 8.636168555094445e-78; dp *= 8.636168555094445e-78; dpm1 *= 8.636168555094445e-78;
        } else if (m < 8.636168555094445e-78) {
            p *= 1.157920892373162e77; pm1 *= 1.157920892373162e77; dp *= 1.157920892373162e77; dpm1 *= 1.157920892373162e77;
        }
    }
    pn_out = p;
    dpn_out = dp;
    return cnt;
}

// Fast form of the above (same treatment of zeros and the same sign bookkeeping as sturm_count).
__device__ __forceinline__ int sturm_newton(const double2* __restrict__ de, int n, double x, double& pn_out, double& dpn_out) {
    double pm1 = 1.0, p = de[0].x - x, dpm1 = 0.0, dp = -1.0;
    unsigned hist = (unsigned)__double2hiint(p) >> 31;
    int cnt = (int)hist;
    int i = 1;
    for (; i + STURM_GROUP <= n; i += STURM_GROUP) {
        hist &= 1u;
#pragma unroll
        for (int k = 0; k < STURM_GROUP; ++k) {
            const double2 q = de[i + k];
            const double t = q.x - x;
            const double pn = fma(t, p, -(q.y * pm1));
            const double dpn = fma(t, dp, -fma(q.y, dpm1, p));
            hist = __funnelshift_l((unsigned)__double2hiint(pn), hist, 1);
            pm1 = p; p = pn;
            dpm1 = dp; dp = dpn;
        }
        cnt += __popc((hist ^ (hist >> 1)) & ((1u << STURM_GROUP) - 1u));
        const double mp = fmax(fabs(p), fabs(pm1));
        if (mp < STURM_TINY) return sturm_newton_guarded(de, n, x, pn_out, dpn_out);
        const double m = fmax(mp, fmax(fabs(dp), fabs(dpm1)) * 0x1p-60);
        if (m > 1.157920892373162e77) {
            p *= 8.636168555094445e-78; pm1 *= 8.636168555094445e-78; dp *= 8.636168555094445e-78; dpm1 *= 8.636168555094445e-78;
        } else if (m < 8.636168555094445e-78) {
            p *= 1.157920892373162e77; pm1 *= 1.157920892373162e77; dp *= 1.157920892373162e77; dpm1 *= 1.157920892373162e77;
        }
    }
    if (i < n) {  // last, shorter group
        const int r = n - i;
        hist &= 1u;
        for (; i < n; ++i) {
            const double2 q = de[i];
            const double t = q.x - x;
            const double pn = fma(t, p, -(q.y * pm1));
            const double dpn = fma(t, dp, -fma(q.y, dpm1, p));
            hist = __funnelshift_l((unsigned)__double2hiint(pn), hist, 1);
            pm1 = p; p = pn;
            dpm1 = dp; dp = dpn;
        }
        cnt += __popc((hist ^ (hist >> 1)) & ((1u << r) - 1u));
        if (fmax(fabs(p), fabs(pm1)) < STURM_TINY) return sturm_newton_guarded(de, n, x, pn_out, dpn_out);
    }
    pn_out = p;
    dpn_out = dp;
    return cnt;
}

// shared-memory block: de[Np] (double2), red[40]
__global__ void __launch_bounds__(1024)
tridiag_eig_kernel(const double* __restrict__ d_all, const double* __restrict__ e_all, int N, double beta,
                   double* __restrict__ evals_all, long evals_stride, const int32_t* __restrict__ slot, long slot_stride,
                   double* __restrict__ out_all, double* __restrict__ exp_all, double* __restrict__ fermi_all,
                   int* __restrict__ flag) {
    extern __shared__ double2 sm2[];
    double2* de = sm2;
    double* red = reinterpret_cast<double*>(sm2 + N);
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const double* d = d_all + (size_t)b * N;
    const double* e = e_all + (size_t)b * N;
    // scale by max |entry| (as Eigen's solver does) so the product-form recurrence stays in range
    double mx = 0.0;
    for (int i = tid; i < N; i += T) mx = fmax(mx, fmax(fabs(d[i]), i < N - 1 ? fabs(e[i]) : 0.0));
    mx = warp_max(mx);
    const int lane_ = tid & 31, warp_ = tid >> 5, nw_ = (T + 31) >> 5;
    if (lane_ == 0) red[warp_] = mx;
    __syncthreads();
    mx = 0.0;
    for (int w = 0; w < nw_; ++w) mx = fmax(mx, red[w]);
    __syncthreads();
    const double sc = mx > 0.0 ? mx : 1.0, isc = 1.0 / sc;
    // Gershgorin bounds of the scaled matrix
    double lo = DBL_MAX, hi = -DBL_MAX;
    for (int i = tid; i < N; i += T) {
        const double el = (i > 0 ? e[i - 1] : 0.0) * isc, er = (i < N - 1 ? e[i] : 0.0) * isc, di = d[i] * isc;
        de[i] = make_double2(di, el * el);
        const double rad = fabs(el) + fabs(er);
        lo = fmin(lo, di - rad);
        hi = fmax(hi, di + rad);
    }
    lo = -warp_max(-lo);
    hi = warp_max(hi);
    {
        if (lane_ == 0) { red[warp_] = lo; red[40 + warp_] = hi; }
        __syncthreads();
        double l2 = DBL_MAX, h2 = -DBL_MAX;
        for (int w = 0; w < nw_; ++w) { l2 = fmin(l2, red[w]); h2 = fmax(h2, red[40 + w]); }
        lo = l2; hi = h2;
        __syncthreads();
    }
    const double span = fmax(hi - lo, DBL_MIN);
    lo -= 4.0 * DBL_EPSILON * span + 2.0 * DBL_MIN;
    hi += 4.0 * DBL_EPSILON * span + 2.0 * DBL_MIN;
    const double scale = fmax(fabs(lo), fabs(hi));

    double* ev = evals_all + (size_t)b * evals_stride + (slot ? (size_t)slot[b] * slot_stride : 0);
    // One cooperative multisection round: thread t counts the eigenvalues below its own grid point, which gives every
    // eigenvalue a bracket of width (hi - lo) / (T + 1) for the price of one Sturm count (instead of ~log2 T bisection steps).
    int* cnts = reinterpret_cast<int*>(red + 96);
    const double gstep = (hi - lo) / (double)(T + 1);
    cnts[tid] = sturm_count(de, N, lo + gstep * (double)(tid + 1));
    __syncthreads();
    double lam = 0.0;
    for (int k = tid; k < N; k += T) {  // T >= N in practice: one eigenvalue per thread
        // first grid index j with cnts[j] > k (counts are non-decreasing); the eigenvalue lies in (x_{j-1}, x_j]
        int jl = 0, jh = T;  // search in [0, T]; index T stands for hi (count N > k)
        while (jl < jh) {
            const int jm = (jl + jh) >> 1;
            if (cnts[jm] > k) jh = jm; else jl = jm + 1;
        }
        double a = (jl == 0) ? lo : lo + gstep * (double)jl;
        double c = (jl >= T) ? hi : lo + gstep * (double)(jl + 1);
        int ca = (jl == 0) ? 0 : cnts[jl - 1];   // eigenvalues below a (<= k)
        int cc = (jl >= T) ? N : cnts[jl];        // eigenvalues below c (> k)
        const double tolw = 2.0 * DBL_EPSILON * scale;
        int it = 0;
        // (1) a fixed number of bisection steps for every lane (keeps the warp in lockstep on the cheap count-only
        //     recurrence and leaves all but ~0.03 % of the eigenvalues isolated), then more only while not isolated
        for (; it < 128 && (it < FKMC_EIG_BISECT || cc - ca > 1); ++it) {
            const double mid = 0.5 * (a + c);
            if (mid <= a || mid >= c || c - a <= tolw) break;
            const int cm = sturm_count(de, N, mid);
            if (cm > k) { c = mid; cc = cm; } else { a = mid; ca = cm; }
        }
        // (2) safeguarded Newton on the characteristic polynomial: every evaluation also shrinks the bracket through its
        //     Sturm count, a step that leaves the bracket (or a non-finite one) is replaced by bisection
        double x = 0.5 * (a + c);
        bool done = !(c - a > tolw);
        int nnewt = 0;
        double prev_step = DBL_MAX;
        for (; it < 128 && !done; ++it) {
            double pv, dpv;
            const int cm = sturm_newton(de, N, x, pv, dpv);
            if (cm > k) c = x; else a = x;
            const double dn = -pv / dpv;
            double xn = x + dn;
            const bool isolated = (cc - ca == 1);
            // a correction at working precision means x already is the eigenvalue (the count may put it on either side of x)
            if (isolated && fabs(dn) <= 4.0 * tolw) {
                a = x; c = x;
                done = true;
                break;
            }
            // Newton only for an isolated root, inside the bracket (also rejects NaN); after 8 Newton steps every other step
            // is a bisection so that an ill-conditioned root cannot stall the iteration
            const bool ok = isolated && (xn > a) && (xn < c) && (nnewt < 8 || (nnewt & 1));
            ++nnewt;
            if (!ok) xn = 0.5 * (a + c);
            // also stop when the correction has stopped contracting (rounding floor of p/p' near the root)
            const double stepn = fabs(xn - x);
            const bool floor_hit = ok && nnewt >= 3 && stepn <= 64.0 * tolw && stepn >= 0.25 * prev_step;
#if FKMC_EIG_PREDSTOP
            // two Newton steps in a row that contract at least 16-fold: e_{k+1} = K e_k^2 with K = s_k / s_{k-1}^2 read off the steps, so
            // the error left after this step is s_k^3 / s_{k-1}^2; below a quarter of the tolerance the confirming evaluation is skipped
            const bool pred_hit = ok && prev_step < 0.5 * DBL_MAX && stepn <= 0.0625 * prev_step &&
                                  stepn * stepn * stepn <= 0.25 * tolw * prev_step * prev_step;
#else
            const bool pred_hit = false;
#endif
            if (floor_hit || pred_hit || !(c - a > tolw) || xn <= a || xn >= c) {
                if (ok) { a = xn; c = xn; }
                done = true;
            }
            prev_step = ok ? stepn : DBL_MAX;
            x = xn;
        }
        if (it >= 128) atomicOr(flag, 1);
#ifdef FKMC_TRIDIAG_DEBUG
        if (fermi_all) fermi_all[(size_t)b * N + k] = (double)(it * 1000 + nnewt);
#endif
        lam = 0.5 * (a + c) * sc;
        ev[k] = lam;
    }
    // ---- fused free energy / Fermi caches / energy measure ----
    __shared__ double e0s;
    if (tid == 0) e0s = ev[0];  // the smallest eigenvalue (thread 0 wrote it itself: k = 0 is its first eigenvalue)
    __syncthreads();
    const double e0 = e0s;
    double lz = 0.0, ec = 0.0, d2 = 0.0;
    for (int k = tid; k < N; k += T) {
        const double x = (T >= N) ? lam : ev[k];
        const double logw0 = beta * e0;
        const double w = exp(-beta * (x - e0));
        const double ex = exp(beta * x);
        lz += log(exp(logw0) + w) - logw0;
        ec += x / (1.0 + ex);
        d2 += x * x / (1.0 + 0.5 * (ex + 1.0 / ex));
        if (exp_all) exp_all[(size_t)b * N + k] = ex;
#ifndef FKMC_TRIDIAG_DEBUG
        if (fermi_all) fermi_all[(size_t)b * N + k] = 1.0 / (1.0 + ex);
#endif
    }
    lz = block_sum(lz, red);
    ec = block_sum(ec, red);
    d2 = block_sum(d2, red);
    if (tid == 0) {
        out_all[(size_t)b * 8 + 0] = lz;
        out_all[(size_t)b * 8 + 1] = ec;
        out_all[(size_t)b * 8 + 2] = 0.5 * d2;
    }
}

// energy measure from an existing spectrum
__global__ void __launch_bounds__(1024)
energy_kernel(const double* __restrict__ evals_all, long evals_stride, const int32_t* __restrict__ slot, long slot_stride, int N,
              double beta, double* __restrict__ out_all) {
    __shared__ double red[40];
    const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
    const double* ev = evals_all + (size_t)b * evals_stride + (slot ? (size_t)slot[b] * slot_stride : 0);
    const double e0 = ev[0];
    double lz = 0.0, ec = 0.0, d2 = 0.0;
    for (int k = tid; k < N; k += T) {
        const double x = ev[k];
        const double logw0 = beta * e0;
        const double w = exp(-beta * (x - e0));
        const double ex = exp(beta * x);
        lz += log(exp(logw0) + w) - logw0;
        ec += x / (1.0 + ex);
        d2 += x * x / (1.0 + 0.5 * (ex + 1.0 / ex));
    }
    lz = block_sum(lz, red);
    ec = block_sum(ec, red);
    d2 = block_sum(d2, red);
    if (tid == 0) {
        out_all[(size_t)b * 8 + 0] = lz;
        out_all[(size_t)b * 8 + 1] = ec;
        out_all[(size_t)b * 8 + 2] = 0.5 * d2;
    }
}

}  // namespace

int fkmc_launch_tridiag_eig(fkmc_ctx* ctx, const double* d_d, const double* d_e, int N, int B, double beta, double* d_evals,
                            long evals_stride, const int32_t* d_slot, long slot_stride, double* d_out, double* d_exp,
                            double* d_fermi) {
    fkmc_prof_scope ps(ctx, "tridiag_eig");
    if (N > 8192) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "tridiag_eig: N > 8192 not supported");
    const int T = std::min(1024, ((N + 31) / 32) * 32);  // one eigenvalue per thread up to N = 1024, several beyond
    const size_t smem = sizeof(double2) * N + sizeof(double) * 96 + sizeof(int) * T + 16;
    tridiag_eig_kernel<<<B, T, smem, ctx->stream>>>(d_d, d_e, N, beta, d_evals, evals_stride, d_slot, slot_stride, d_out, d_exp,
                                                    d_fermi, ctx->d_flag);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

int fkmc_launch_energy(fkmc_ctx* ctx, const double* d_evals, long evals_stride, const int32_t* d_slot, long slot_stride, int N,
                       int B, double beta, double* d_out) {
    fkmc_prof_scope ps(ctx, "energy");
    int T = ((N + 31) / 32) * 32;
    if (T > 1024) T = 1024;
    energy_kernel<<<B, T, 0, ctx->stream>>>(d_evals, evals_stride, d_slot, slot_stride, N, beta, d_out);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}
