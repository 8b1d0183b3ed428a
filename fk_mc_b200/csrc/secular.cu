// Fast re-weighting of the dense (exact-diagonalisation) moves: a proposal changes one (add_remove) or two (flip) diagonal
// entries of H, so with the eigen-decomposition H = V diag(lam) V^T of the CURRENT configuration at hand the new spectrum is the
// root set of a rank-one secular equation -- O(N^2) per proposal instead of the 4/3 N^3 of a fresh eigensolve -- and only an
// ACCEPTED move pays an N^3 update of the eigenvectors (V <- V Q, a DMMA GEMM whose Cauchy-like right factor is generated on the fly).
//
// This is SURVEY 8(f)-3: it replaces new_config.calc_ed(false) inside move_addremove::attempt / move_flip::attempt
// (src/moves.cpp:5-21,52-67; the reference's own benchmark of this step is benchmark/fast_update.cpp) without changing what is
// computed: eigenvalues agree with the full solve to ~1e-13, the accept/reject sequence is the one of the full-solve path.
//
//   H' = H + rho e_i e_i^T,  z = V^T e_i (row i of V)   =>   eig(H') = eig(diag(lam) + rho z z^T):
//   roots of  f(x) = 1 + rho sum_k z_k^2 / (lam_k - x),  one in every gap of the poles (strict interlacing).
// Root j is found in the variable mu = x - lam_o with o the nearer pole (origin), so that all pole distances
// (lam_k - lam_o) - mu keep high relative accuracy; the iteration is the two-pole rational interpolation of
// Bunch-Nielsen-Sorensen (monotone, quadratically convergent) safeguarded by the bracket.  The eigenvectors of the updated
// matrix follow Gu & Eisenstat: zhat_k^2 = prod_j (lam'_j - lam_k) / (rho prod_{j != k} (lam_j - lam_k)) is recomputed from the
// computed roots, which makes Q[k][j] = zhat_k / ((lam_k - lam'_j) nrm_j) orthogonal to working precision whatever the gaps.
// No deflation logic: vanishing components z_k and coinciding poles are lifted to a floor (a perturbation of H of order 1e-15).
// A flip is two successive rank-one steps (remove at `from`, add at `to`); z of the second step is Q_1^T (V^T e_to).
//
// Every `fu_refresh` sweeps the eigen-decomposition is recomputed from scratch (eigvec.cu) and compared with the tracked
// spectrum; a trace invariant is checked on every proposal.  A violation raises bit 4 of the non-convergence flag
// (FKMC_ERR_NOCONV from fkmc_chain_run_sweeps).
#include <cfloat>

#include "common.cuh"
#include "gemm_tile.cuh"

namespace {

constexpr double Z2_FLOOR = 1e-34;  // floor of z_k^2 (|z| = 1): keeps every gap's root strictly inside the gap
constexpr int FU_MAXIT = 60;

struct fu_args {
    int N, n_chains;
    double U, beta;
    const double* spec;       // [2][C][N] spectra, slot-major
    const int32_t* cur_slot;  // [C]
    const int32_t* prop_slot; // [C]
    const double* vt;         // [2][C][N][N] site-major eigenvectors: vt[slot][c][i][k] = component i of eigenvector k
    const int32_t* vslot;     // [C]
    const int32_t *prop_move, *prop_a, *prop_b;
    const int32_t* f_cur;     // [C][V]
    double* poles;            // [2][C][N] poles of stage 0 / 1
    int32_t* org;             // [2][C][N] origin pole of root j
    double* mu;               // [2][C][N] root j = poles[org[j]] + mu[j]
    double* zhat;             // [2][C][N]
    double* inrm;             // [2][C][N] 1 / norm of eigenvector column j
    int32_t* nstage;          // [C]
    double* rho;              // [2][C] strength of stage 0 / 1 of the pending proposal
    double* out;              // [C][8]: logZ, E_c, d2E of the proposal
    int* flag;
};

// Block reductions written WITHOUT lane-0 predicates: nvcc 12.9 (sm_100a, -O3) folds `buf + 8 (tid >> 5)` into `buf + (tid >> 2)` under a
// `(tid & 31) == 0` predicate and then reuses that address for unpredicated per-warp accesses of the same inlined function, which land a few
// bytes off (compute-sanitizer: misaligned shared access).  After a butterfly reduction every lane holds the result, so all lanes store it.
__device__ __forceinline__ double block_sum_t(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    red[warp] = v;
    __syncthreads();
    double s = lane < nw ? red[lane] : 0.0;
    s = warp_sum(s);
    __syncthreads();
    return s;
}
__device__ __forceinline__ double block_max_t(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    red[warp] = v;
    __syncthreads();
    double s = lane < nw ? red[lane] : -DBL_MAX;
    s = warp_max(s);
    __syncthreads();
    return s;
}

// 1 / x to full double precision without the special-case handling of the IEEE division sequence: MUFU.RCP64H seed (2^-23) and two
// Newton steps.  |x| is a distance between a trial root and a pole: never zero, never denormal in practice (floors above).
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// inclusive max-scan over the block (n <= blockDim.x values, one per thread; identity -inf for the rest)
__device__ __forceinline__ double block_max_scan(double v, double* buf) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = fmax(v, u);
    }
    const double vlast = __shfl_sync(0xffffffffu, v, 31);
    buf[warp] = vlast;  // every lane stores the same value (see block_sum_t)
    __syncthreads();
    // every warp scans the per-warp totals itself (nw <= 32)
    double w = lane < nw ? buf[lane] : -DBL_MAX;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w = fmax(w, u);
    }
    const double prev = __shfl_sync(0xffffffffu, w, (warp + 31) & 31);  // inclusive total of the warps before this one
    if (warp > 0) v = fmax(v, prev);
    __syncthreads();
    return v;
}

// One rank-one stage on the poles d[0..N) (ascending, strictly separated, in shared memory) with weights z2[0..N) (> 0) and
// strength rho: thread j < N returns root j (ascending order) as (origin, mu).  sumz2 = sum z2.  Shared arrays cd / cz hold the
// canonical problem (rho > 0): for rho < 0 the poles are negated and reversed.
__device__ __forceinline__ void secular_stage(int N, int j, const double* d, const double* z2, double rho, double sumz2, double* cd, double* cz, int& org_out,
                              double& mu_out, bool& ok) {
    const int tid = threadIdx.x;
    const bool rev = rho < 0.0;
    const double R = fabs(rho);
    if (tid < N) {
        const int s = rev ? N - 1 - tid : tid;
        cd[tid] = rev ? -d[s] : d[s];
        cz[tid] = z2[s];
    }
    __syncthreads();
    ok = true;
    if (j < N) {
        const int jc = rev ? N - 1 - j : j;  // canonical root index of the thread's root
        const bool last = (jc == N - 1);
        const double dl = cd[jc];
        const double gap = last ? R * sumz2 : cd[jc + 1] - dl;  // root in (dl, dl + gap)
        // which end is nearer: sign of f at the midpoint (poles measured from dl)
        int o = jc;
        double m0 = NAN;   // initial guess (in x - dl) from the midpoint pass
        if (!last) {
            const double xm = 0.5 * gap;
            double fm = 0.0;
            for (int k = 0; k < N; ++k) fm = fma(cz[k], fast_rcp((cd[k] - dl) - xm), fm);
            fm = fma(R, fm, 1.0);
            if (fm < 0.0) o = jc + 1;  // root in the upper half
            // dlaed4-style start: the two neighbouring poles exactly, the rest frozen at its midpoint value c:
            //   c - zl / x + zu / (gap - x) = 0   <=>   c x^2 - (c gap + zl + zu) x + zl gap = 0,  x in (0, gap)
            const double zl = R * cz[jc], zu = R * cz[jc + 1];
            const double c = fm - 2.0 * (zu - zl) / gap;
            const double b = -(c * gap + zl + zu), q = zl * gap;
            if (c == 0.0) {
                m0 = -q / b;
            } else {
                const double sq = sqrt(fmax(fma(b, b, -4.0 * c * q), 0.0));
                const double t = (b > 0.0) ? -(b + sq) : (sq - b);
                const double x1 = 0.5 * t / c, x2 = (t != 0.0) ? 2.0 * q / t : 0.0;
                m0 = (x1 > 0.0 && x1 < gap) ? x1 : x2;
            }
        }
        const double dorg = cd[o];
        // bracket in mu = x - dorg
        double lo, hi;
        if (o == jc) { lo = 0.0; hi = last ? gap : 0.5 * gap; }
        else { lo = -0.5 * gap; hi = 0.0; }
        const double dj = cd[jc] - dorg;                       // lower pole in the shifted variable (0 or -gap)
        const double du = last ? 0.0 : cd[jc + 1] - dorg;      // upper pole (unused for the last root)
        double m = (o == jc) ? m0 : m0 - gap;
        if (!(m > lo && m < hi)) m = (o == jc) ? (last ? fmin(0.5 * gap, R * cz[jc]) : 0.25 * gap) : -0.25 * gap;  // fall-back start inside the bracket
        if (!(m > lo && m < hi)) m = 0.5 * (lo + hi);
        bool done = false;
        for (int it = 0; it < FU_MAXIT && !done; ++it) {
            // psi: poles <= jc, phi: poles > jc; values and derivatives
            double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
            for (int k = 0; k < N; ++k) {   // one uniform loop for the whole warp (shared-memory broadcasts); the split point is per thread
                const double r = fast_rcp((cd[k] - dorg) - m);
                const double t = cz[k] * r, t2 = t * r;
                const bool low = k <= jc;
                psi += low ? t : 0.0;
                dpsi += low ? t2 : 0.0;
                phi += low ? 0.0 : t;
                dphi += low ? 0.0 : t2;
            }
            psi *= R; dpsi *= R; phi *= R; dphi *= R;
            const double fv = 1.0 + psi + phi;
            // f is increasing: shrink the bracket
            if (fv > 0.0) hi = fmin(hi, m); else lo = fmax(lo, m);
            if (fv == 0.0) break;
            // rounding level of f: when |f| is below it the root is resolved
            const double ferr = 8.0 * DBL_EPSILON * (1.0 + fabs(psi) + fabs(phi));
            if (fabs(fv) <= ferr) break;   // m is the root to rounding: keep it
            double eta;
            if (last) {
                // one-sided: interpolate psi by s / (dj - x) + p through value and slope -> x+ = dj + s / (1 + p) ... as a correction
                const double Dj = dj - m;
                const double s = Dj * Dj * dpsi, p = psi - Dj * dpsi;
                const double c = 1.0 + p;
                eta = (c > 0.0) ? (Dj + s / c) : (hi - m) * 0.5;   // new mu = dj + s / c  <=>  eta = Dj + s / c  (Dj < 0 < s)
            } else {
                const double Dj = dj - m, Du = du - m;           // Dj < 0 < Du
                const double s = Dj * Dj * dpsi, S = Du * Du * dphi;
                const double c = 1.0 + (psi - Dj * dpsi) + (phi - Du * dphi);
                // c eta^2 + b eta + q = 0 with q = Dj Du f
                const double b = -(c * (Dj + Du) + s + S);
                const double q = Dj * Du * fv;
                const double disc = fma(b, b, -4.0 * c * q);
                if (c == 0.0) {
                    eta = -q / b;
                } else {
                    const double sq = sqrt(fmax(disc, 0.0));
                    // the root of the quadratic inside (Dj, Du): eta has the sign of -f (f increasing)
                    const double t = (b > 0.0) ? -(b + sq) : (sq - b);
                    const double e1 = 0.5 * t / c, e2 = (t != 0.0) ? 2.0 * q / t : 0.0;
                    const bool in1 = e1 > Dj && e1 < Du, in2 = e2 > Dj && e2 < Du;
                    eta = (in1 && in2) ? (fabs(e1) < fabs(e2) ? e1 : e2) : (in1 ? e1 : e2);
                    if (!in1 && !in2) eta = NAN;
                }
            }
            double mn = m + eta;
            const bool interp = (mn > lo && mn < hi);
            if (!interp) mn = 0.5 * (lo + hi);                     // safeguard: bisection
            if (mn == m || hi - lo <= 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi))) done = true;
            // the interpolation converges quadratically: a correction below 1e-9 |mu| leaves an error at rounding level, so it is applied
            // without a confirming pass
            if (interp && fabs(mn - m) <= 1e-9 * fabs(mn)) done = true;
            if (fabs(mn - m) <= 4.0 * DBL_EPSILON * fabs(mn)) done = true;
            m = mn;
            if (it == FU_MAXIT - 1 && !done) ok = false;
        }
        // back to the original frame
        if (rev) { org_out = N - 1 - o; mu_out = -m; }
        else { org_out = o; mu_out = m; }
    }
}

// sites -> proposal spectrum.  One CTA per chain, one thread per root (N <= 1024).
// shared: d[N] z2[N] cd[N] cz[N] lamn[N] zs[N] (signed z) zh[N] red[72]
__global__ void __launch_bounds__(1024) fu_eval_kernel(fu_args P) {
    extern __shared__ __align__(16) double sm[];
    const int N = P.N, c = blockIdx.x, tid = threadIdx.x;
    double* d = sm;
    double* z2 = d + N;
    double* cd = z2 + N;
    double* cz = cd + N;
    double* lamn = cz + N;
    double* zs = lamn + N;
    double* zh = zs + N;
    double* red = zh + N;
    const size_t CN = (size_t)P.n_chains * N;
    const int kind = P.prop_move[c];
    const double* lam = P.spec + (size_t)P.cur_slot[c] * CN + (size_t)c * N;
    double* lam_out = const_cast<double*>(P.spec) + (size_t)P.prop_slot[c] * CN + (size_t)c * N;
    const double* vt = P.vt + ((size_t)P.vslot[c] * P.n_chains + c) * (size_t)N * N;
    int nst = 0, site0 = -1, site1 = -1;
    double rho0 = 0.0, rho1 = 0.0;
    if (kind == FKMC_MOVE_ADDREMOVE) {
        nst = 1;
        site0 = P.prop_a[c];
        rho0 = P.f_cur[(size_t)c * N + site0] ? -P.U : P.U;   // f toggles: 0 -> 1 adds +U on the diagonal
    } else if (kind == FKMC_MOVE_FLIP) {
        nst = 2;
        site0 = P.prop_a[c]; rho0 = -P.U;   // from: occupied -> empty
        site1 = P.prop_b[c]; rho1 = P.U;    // to: empty -> occupied
    }
    if (P.U == 0.0) nst = 0;  // nothing changes: roots == poles
    if (tid == 0) {
        P.nstage[c] = nst;
        P.rho[c] = rho0;
        P.rho[P.n_chains + c] = rho1;
    }
    // scale of the spectrum (separation floor)
    double scale = 0.0;
    if (tid < N) scale = fabs(lam[tid]);
    scale = block_max_t(fmax(scale, 1.0), red);
    const double gmin = 4.0 * DBL_EPSILON * scale;
    double cur = (tid < N) ? lam[tid] : 0.0;   // this thread's pole of the running stage
    bool all_ok = true;
    for (int st = 0; st < nst; ++st) {
        const int site = st == 0 ? site0 : site1;
        const double rho = st == 0 ? rho0 : rho1;
        // poles: strictly separated copy of the running spectrum  d_k = k g + max_{i <= k} (lam_i - i g)
        {
            double v = (tid < N) ? cur - (double)tid * gmin : -DBL_MAX;
            v = block_max_scan(v, red);
            if (tid < N) d[tid] = v + (double)tid * gmin;
        }
        // z = V^T e_site in the basis of this stage
        if (st == 0) {
            if (tid < N) zs[tid] = vt[(size_t)site * N + tid];
            __syncthreads();
        } else {
            // z^1_j = inrm_j sum_k zhat_k V[site][k] / ((d_k - d_o) - mu_j): Q_0^T applied to row `site` of V  (stage-0 data in zh, lamn frame)
            const double* p0 = P.poles + (size_t)c * N;   // stage-0 poles as stored
            const double* zh0 = P.zhat + (size_t)c * N;
            double acc = 0.0;
            if (tid < N) {
                const int o = P.org[(size_t)c * N + tid];
                const double m = P.mu[(size_t)c * N + tid], po = p0[o];
                for (int k = 0; k < N; ++k) acc = fma(zh0[k] * vt[(size_t)site * N + k], 1.0 / ((p0[k] - po) - m), acc);
                acc *= P.inrm[(size_t)c * N + tid];
            }
            __syncthreads();
            if (tid < N) zs[tid] = acc;
            __syncthreads();
        }
        double zz = 0.0;
        if (tid < N) {
            zz = fmax(zs[tid] * zs[tid], Z2_FLOOR);
            z2[tid] = zz;
        }
        const double sumz2 = block_sum_t(zz, red);
        int o = 0;
        double m = 0.0;
        bool ok = true;
        secular_stage(N, tid, d, z2, rho, sumz2, cd, cz, o, m, ok);
        all_ok = all_ok && ok;
        const size_t base = ((size_t)st * P.n_chains + c) * N;
        double ln = 0.0;
        if (tid < N) {
            ln = d[o] + m;
            lamn[tid] = ln;
            P.poles[base + tid] = d[tid];
            P.org[base + tid] = o;
            P.mu[base + tid] = m;
            if (st + 1 == nst) P.zhat[base + tid] = copysign(1.0, zs[tid]);  // sign(z_k) for fu_prepare_kernel (runs only if accepted)
        }
        // trace invariant: sum lam' = sum d + rho sum z2
        {
            const double tr = block_sum_t((tid < N) ? (ln - d[tid]) : 0.0, red);
            if (tid == 0 && fabs(tr - rho * sumz2) > 1e-9 * scale) atomicOr(P.flag, 4);
        }
        if (st + 1 < nst) {
            // Gu-Eisenstat zhat and the column norms of stage 0 (needed for z of the second stage); lamn as (origin, mu) pairs in global
            __syncthreads();
            __threadfence_block();
            const int32_t* og = P.org + base;
            const double* mg = P.mu + base;
            double zk = 0.0;
            if (tid < N) {
                const double dk = d[tid];
                double prod = ((d[og[tid]] - dk) + mg[tid]) / rho;
                for (int jj = 0; jj < N; ++jj) {
                    if (jj == tid) continue;
                    prod *= ((d[og[jj]] - dk) + mg[jj]) / (d[jj] - dk);
                }
                zk = copysign(sqrt(fmax(prod, 0.0)), zs[tid]);
                zh[tid] = zk;
                P.zhat[base + tid] = zk;
            }
            __syncthreads();
            if (tid < N) {
                const double po = d[og[tid]], mm = mg[tid];
                double s = 0.0;
                for (int k = 0; k < N; ++k) {
                    const double r = zh[k] / ((d[k] - po) - mm);
                    s = fma(r, r, s);
                }
                P.inrm[base + tid] = rsqrt(s);
            }
            __syncthreads();
            __threadfence_block();
        }
        cur = ln;
        __syncthreads();
    }
    if (nst == 0) {
        // "this move won't work" (weight 0, always rejected): the proposal slot just mirrors the current spectrum
        if (tid < N) lamn[tid] = cur;
        __syncthreads();
    }
    if (!all_ok) atomicOr(P.flag, 1);
    // proposal spectrum + fused free energy / energy measure (configuration.cpp:226-244, measures/energy.cpp:6-26)
    if (tid < N) lam_out[tid] = lamn[tid];
    const double e0 = lamn[0];
    double lz = 0.0, ec = 0.0, d2 = 0.0;
    if (tid < N) {
        const double x = lamn[tid];
        const double logw0 = P.beta * e0;
        const double w = exp(-P.beta * (x - e0));
        const double ex = exp(P.beta * x);
        lz = log(exp(logw0) + w) - logw0;
        ec = x / (1.0 + ex);
        d2 = x * x / (1.0 + 0.5 * (ex + 1.0 / ex));
    }
    lz = block_sum_t(lz, red);
    ec = block_sum_t(ec, red);
    d2 = block_sum_t(d2, red);
    if (tid == 0) {
        P.out[(size_t)c * 8 + 0] = lz;
        P.out[(size_t)c * 8 + 1] = ec;
        P.out[(size_t)c * 8 + 2] = 0.5 * d2;
    }
}

// stage-level entry (tests): roots of diag(lam) + rho z z^T for a batch; one CTA per problem
__global__ void __launch_bounds__(1024) secular_only_kernel(int N, const double* __restrict__ lam_all, const double* __restrict__ z_all, const double* __restrict__ rho_all,
                                                          double* __restrict__ out_all, int* flag) {
    extern __shared__ __align__(16) double sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    double* d = sm;
    double* z2 = d + N;
    double* cd = z2 + N;
    double* cz = cd + N;
    double* red = cz + N;
    const double* lam = lam_all + (size_t)b * N;
    const double rho = rho_all[b];
    double scale = (tid < N) ? fmax(fabs(lam[tid]), 1.0) : 1.0;
    scale = block_max_t(scale, red);
    const double gmin = 4.0 * DBL_EPSILON * scale;
    double v = (tid < N) ? lam[tid] - (double)tid * gmin : -DBL_MAX;
    v = block_max_scan(v, red);
    double zz = 0.0;
    if (tid < N) {
        d[tid] = v + (double)tid * gmin;
        const double zv = z_all[(size_t)b * N + tid];
        zz = fmax(zv * zv, Z2_FLOOR);
        z2[tid] = zz;
    }
    const double sumz2 = block_sum_t(zz, red);
    int o = 0;
    double m = 0.0;
    bool ok = true;
    if (rho != 0.0) secular_stage(N, tid, d, z2, rho, sumz2, cd, cz, o, m, ok);
    else o = tid < N ? tid : 0;
    if (!ok) atomicOr(flag, 1);
    if (tid < N) out_all[(size_t)b * N + tid] = d[o] + m;
}

// accepted chains: zhat / column norms of the last stage (stage 0 of a flip already has them)
__global__ void __launch_bounds__(1024) fu_prepare_kernel(fu_args P, const int32_t* __restrict__ accepted) {
    extern __shared__ __align__(16) double sm[];
    const int N = P.N, c = blockIdx.x, tid = threadIdx.x;
    if (!accepted[c]) return;
    const int nst = P.nstage[c];
    if (nst == 0) return;
    const int st = nst - 1;
    double* d = sm;
    double* lamd = d + N;   // (d[org[j]] , mu[j]) resolved
    double* zh = lamd + N;
    double* mus = zh + N;
    const size_t base = ((size_t)st * P.n_chains + c) * N;
    const double rho = P.rho[(size_t)st * P.n_chains + c];
    if (tid < N) d[tid] = P.poles[base + tid];
    __syncthreads();
    if (tid < N) {
        lamd[tid] = d[P.org[base + tid]];
        mus[tid] = P.mu[base + tid];
    }
    __syncthreads();
    // the evaluation kernel left sign(z_k) of the last stage in zhat (as +-1)
    if (tid < N) {
        const double dk = d[tid];
        double prod = ((lamd[tid] - dk) + mus[tid]) / rho;
        for (int jj = 0; jj < N; ++jj) {
            if (jj == tid) continue;
            prod *= ((lamd[jj] - dk) + mus[jj]) / (d[jj] - dk);
        }
        const double sgn = P.zhat[base + tid];  // +-1 left by the evaluation kernel
        const double zk = copysign(sqrt(fmax(prod, 0.0)), sgn);
        zh[tid] = zk;
    }
    __syncthreads();
    if (tid < N) {
        P.zhat[base + tid] = zh[tid];
        const double po = lamd[tid], mm = mus[tid];
        double s = 0.0;
        for (int k = 0; k < N; ++k) {
            const double r = zh[k] / ((d[k] - po) - mm);
            s = fma(r, r, s);
        }
        P.inrm[base + tid] = rsqrt(s);
    }
}

// ---- V <- V Q for accepted chains: C[i][j] = sum_k A[i][k] Q[k][j],  Q[k][j] = zhat_k inrm_j / ((p_k - p_{o_j}) - mu_j) ----
struct gemm_args {
    int N, n_chains, stage;
    double* vt;                // [2][C][N][N]
    double* q;                 // [C][N][N] right factor of the accepted chains, row-major Q[k][j]
    const int32_t* vslot;
    const int32_t* accepted;
    const int32_t* nstage;
    const double *poles, *mu, *zhat, *inrm;  // [2][C][N]
    const int32_t* org;
};

// Q[k][j] = zhat_k inrm_j / ((p_k - p_{o_j}) - mu_j) of one stage, written once per accepted chain (N^2 divisions), so that the
// N^3 product below is a plain GEMM whose inner loop carries no FP64 divisions competing with the DMMAs for the FP64 pipe.
// grid (ceil(N / 32), C): one CTA fills 32 rows.
__global__ void __launch_bounds__(256) fu_qgen_kernel(gemm_args P) {
    const int c = blockIdx.y;
    if (!P.accepted[c] || P.stage >= P.nstage[c]) return;
    const int N = P.N, tid = threadIdx.x;
    const size_t base = ((size_t)P.stage * P.n_chains + c) * N;
    __shared__ double pk[32], zk[32];
    const int k0 = blockIdx.x * 32;
    if (tid < 32 && k0 + tid < N) {
        pk[tid] = P.poles[base + k0 + tid];
        zk[tid] = P.zhat[base + k0 + tid];
    }
    __syncthreads();
    double* Q = P.q + (size_t)c * N * N;
    for (int j = tid; j < N; j += 256) {
        const double pj = P.poles[base + P.org[base + j]], mj = P.mu[base + j], nj = P.inrm[base + j];
#pragma unroll 8
        for (int r = 0; r < 32; ++r)
            if (k0 + r < N) Q[(size_t)(k0 + r) * N + j] = zk[r] * nj / ((pk[r] - pj) - mj);
    }
}

// C = A Q for accepted chains: A = eigenvectors (site-major, [i][k]) in the source slot, C -> the other slot (gemm_tile.cuh).
__global__ void __launch_bounds__(256, 2) fu_gemm_kernel(gemm_args P) {
    const int c = blockIdx.y;
    if (!P.accepted[c] || P.stage >= P.nstage[c]) return;
    extern __shared__ __align__(16) double sm[];
    // stage 0 reads slot s and writes slot 1 - s; stage 1 (second half of a flip) goes back
    const int s0 = P.vslot[c];
    const int src = P.stage == 0 ? s0 : 1 - s0, dst = 1 - src;
    const size_t NN = (size_t)P.N * P.N;
    fkgemm::tile(P.vt + ((size_t)src * P.n_chains + c) * NN, P.q + (size_t)c * NN, P.vt + ((size_t)dst * P.n_chains + c) * NN, P.N, blockIdx.x, sm);
}

__global__ void fu_commit_kernel(int n_chains, const int32_t* __restrict__ accepted, const int32_t* __restrict__ nstage, int32_t* __restrict__ vslot) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    if (accepted[c] && nstage[c] == 1) vslot[c] ^= 1;  // one GEMM moved the eigenvectors to the other slot; a flip's two GEMMs come back
}

// measure_ipr (include/fk_mc/measures/ipr.hpp:39-56) from the tracked eigenvectors: ipr_k = ||psi_k||_4 / ||psi_k||_2^2.  thread = state k
__global__ void __launch_bounds__(256) fu_ipr_kernel(int N, int n_chains, const double* __restrict__ vt_all, const int32_t* __restrict__ vslot, double* __restrict__ ipr) {
    const int c = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
    if (k >= N) return;
    const double* vt = vt_all + ((size_t)vslot[c] * n_chains + c) * (size_t)N * N;
    double s2 = 0.0, s4 = 0.0;
    for (int i = 0; i < N; ++i) {
        const double x = vt[(size_t)i * N + k], x2 = x * x;
        s2 += x2;
        s4 = fma(x2, x2, s4);
    }
    ipr[(size_t)c * N + k] = sqrt(sqrt(s4)) / s2;
}

// refresh: scatter the freshly computed spectra into the chains' current slots and compare with the tracked ones
__global__ void __launch_bounds__(256) fu_refresh_kernel(int N, int n_chains, const double* __restrict__ fresh, double* __restrict__ spec, const int32_t* __restrict__ cur_slot,
                                                        int32_t* __restrict__ vslot, double tol, int check, int* flag, double* __restrict__ maxdev) {
    const int c = blockIdx.x;
    double* cur = spec + (size_t)cur_slot[c] * n_chains * N + (size_t)c * N;
    double dev = 0.0, sc = 1.0;
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        const double v = fresh[(size_t)c * N + k];
        dev = fmax(dev, fabs(v - cur[k]));
        sc = fmax(sc, fabs(v));
        cur[k] = v;
    }
    __shared__ double red[40];
    dev = warp_max(dev); sc = warp_max(sc);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = dev; red[8 + (threadIdx.x >> 5)] = sc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { dev = fmax(dev, red[w]); sc = fmax(sc, red[8 + w]); }
        if (check && dev > tol * sc) atomicOr(flag, 4);
        if (maxdev) maxdev[c] = dev / sc;
        vslot[c] = 0;  // the fresh eigenvectors were written to slot 0
    }
}

}  // namespace

static size_t fu_eval_smem(int N) { return sizeof(double) * (7 * (size_t)N + 80); }

int fkmc_fu_alloc(fkmc_ctx* ctx) {
    fkmc_chain_state& S = ctx->chain;
    const size_t C = S.n_chains, N = ctx->N;
    if (N > 1024) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "fast_update needs N <= 1024");
    auto al = [&](void** p, size_t bytes) { return cudaMalloc(p, bytes) == cudaSuccess ? 0 : 1; };
    int rc = 0;
    rc |= al((void**)&S.fu_vt, sizeof(double) * 2 * C * N * N);
    rc |= al((void**)&S.fu_q, sizeof(double) * C * N * N);
    rc |= al((void**)&S.fu_vslot, sizeof(int32_t) * C);
    rc |= al((void**)&S.fu_poles, sizeof(double) * 2 * C * N);
    rc |= al((void**)&S.fu_org, sizeof(int32_t) * 2 * C * N);
    rc |= al((void**)&S.fu_mu, sizeof(double) * 2 * C * N);
    rc |= al((void**)&S.fu_zhat, sizeof(double) * 2 * C * N);
    rc |= al((void**)&S.fu_inrm, sizeof(double) * 2 * C * N);
    rc |= al((void**)&S.fu_nstage, sizeof(int32_t) * C);
    rc |= al((void**)&S.fu_rho, sizeof(double) * 2 * C);
    rc |= al((void**)&S.fu_acc, sizeof(int32_t) * C);
    rc |= al((void**)&S.fu_fresh, sizeof(double) * C * N);
    rc |= al((void**)&S.fu_maxdev, sizeof(double) * C);
    if (rc) {
        cudaGetLastError();
        fkmc_fu_free(ctx);   // whatever was allocated before the failure
        return fkmc_set_error(ctx, FKMC_ERR_CUDA, "fast_update: out of device memory (3 N^2 doubles per chain)");
    }
    FKMC_CUDA(ctx, cudaMemsetAsync(S.fu_vslot, 0, sizeof(int32_t) * C, ctx->stream));
    FKMC_CUDA(ctx, cudaMemsetAsync(S.fu_acc, 0, sizeof(int32_t) * C, ctx->stream));
    FKMC_CUDA(ctx, cudaMemsetAsync(S.fu_nstage, 0, sizeof(int32_t) * C, ctx->stream));
    return FKMC_OK;
}

void fkmc_fu_free(fkmc_ctx* ctx) {
    fkmc_chain_state& S = ctx->chain;
    cudaFree(S.fu_vt); cudaFree(S.fu_q); cudaFree(S.fu_vslot); cudaFree(S.fu_poles); cudaFree(S.fu_org); cudaFree(S.fu_mu); cudaFree(S.fu_zhat);
    cudaFree(S.fu_inrm); cudaFree(S.fu_nstage); cudaFree(S.fu_rho); cudaFree(S.fu_acc); cudaFree(S.fu_fresh); cudaFree(S.fu_maxdev);
    S.fu_vt = S.fu_q = S.fu_poles = S.fu_mu = S.fu_zhat = S.fu_inrm = S.fu_rho = S.fu_fresh = S.fu_maxdev = nullptr;
    S.fu_vslot = S.fu_org = S.fu_nstage = S.fu_acc = nullptr;
}

static fu_args make_fu_args(fkmc_ctx* ctx) {
    fkmc_chain_state& S = ctx->chain;
    fu_args P{};
    P.N = ctx->N; P.n_chains = S.n_chains; P.U = S.p.U; P.beta = S.p.beta;
    P.spec = S.spec[0]; P.cur_slot = S.cur_slot; P.prop_slot = S.prop_slot; P.vt = S.fu_vt; P.vslot = S.fu_vslot;
    P.prop_move = S.prop_move; P.prop_a = S.prop_a; P.prop_b = S.prop_b; P.f_cur = S.f_cur;
    P.poles = S.fu_poles; P.org = S.fu_org; P.mu = S.fu_mu; P.zhat = S.fu_zhat; P.inrm = S.fu_inrm; P.nstage = S.fu_nstage; P.rho = S.fu_rho;
    P.out = ctx->d_out; P.flag = ctx->d_flag;
    return P;
}

// full eigen-decomposition of the current configurations -> eigenvectors (slot 0) and spectrum (current slot); check != 0 compares
// the tracked spectrum with the fresh one first
int fkmc_fu_refresh(fkmc_ctx* ctx, int check) {
    fkmc_chain_state& S = ctx->chain;
    const int C = S.n_chains, N = ctx->N;
    int rc = fkmc_eigvec_pipeline_dev(ctx, S.f_cur, C, S.p.U, S.p.mu_c, S.p.beta, S.fu_fresh, ctx->d_out, S.fu_vt);
    if (rc) return rc;
    fkmc_prof_scope ps(ctx, "fu_refresh");
    fu_refresh_kernel<<<C, 256, 0, ctx->stream>>>(N, C, S.fu_fresh, S.spec[0], S.cur_slot, S.fu_vslot, 1e-10, check, ctx->d_flag, S.fu_maxdev);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

// proposal spectra by secular updates: fills the proposal slots of the spectra and d_out[c][0..2]
int fkmc_fu_evaluate(fkmc_ctx* ctx) {
    fkmc_chain_state& S = ctx->chain;
    const int N = ctx->N, T = ((N + 31) / 32) * 32;
    fu_args P = make_fu_args(ctx);
    fkmc_prof_scope ps(ctx, "fu_eval");
    const size_t smem = fu_eval_smem(N);
    FKMC_CUDA(ctx, cudaFuncSetAttribute(fu_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fu_eval_kernel<<<S.n_chains, T, smem, ctx->stream>>>(P);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

// after the accept kernel (S.fu_acc[c] = 1 for accepted chains): V <- V Q_0 (Q_1)
int fkmc_fu_commit(fkmc_ctx* ctx) {
    fkmc_chain_state& S = ctx->chain;
    const int N = ctx->N, C = S.n_chains, T = ((N + 31) / 32) * 32;
    fu_args P = make_fu_args(ctx);
    {
        fkmc_prof_scope ps(ctx, "fu_prepare");
        fu_prepare_kernel<<<C, T, sizeof(double) * 4 * N, ctx->stream>>>(P, S.fu_acc);
        ctx->launches++;
    }
    gemm_args G{};
    G.N = N; G.n_chains = C; G.vt = S.fu_vt; G.q = S.fu_q; G.vslot = S.fu_vslot; G.accepted = S.fu_acc; G.nstage = S.fu_nstage;
    G.poles = S.fu_poles; G.mu = S.fu_mu; G.zhat = S.fu_zhat; G.inrm = S.fu_inrm; G.org = S.fu_org;
    const size_t smem = fkgemm::smem_bytes();
    FKMC_CUDA(ctx, cudaFuncSetAttribute(fu_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = fkgemm::tiles(N);
    const int nstages = (S.p.mc_flip > 0.0) ? 2 : 1;
    {
        fkmc_prof_scope ps(ctx, "fu_gemm");
        for (int st = 0; st < nstages; ++st) {
            G.stage = st;
            fu_qgen_kernel<<<dim3((N + 31) / 32, C), 256, 0, ctx->stream>>>(G);
            fu_gemm_kernel<<<dim3(tiles, C), 256, smem, ctx->stream>>>(G);
            ctx->launches += 2;
        }
    }
    fu_commit_kernel<<<(C + 127) / 128, 128, 0, ctx->stream>>>(C, S.fu_acc, S.fu_nstage, S.fu_vslot);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

int fkmc_launch_secular_only(fkmc_ctx* ctx, int N, int B, const double* d_lam, const double* d_z, const double* d_rho, double* d_out) {
    if (N > 1024) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "secular update: N > 1024");
    const int T = ((N + 31) / 32) * 32;
    secular_only_kernel<<<B, T, sizeof(double) * (4 * (size_t)N + 80), ctx->stream>>>(N, d_lam, d_z, d_rho, d_out, ctx->d_flag);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}

// IPR of the chains' current eigenstates from the tracked eigenvectors (no eigensolve): d_ipr [n_chains][N]
int fkmc_fu_ipr(fkmc_ctx* ctx, double* d_ipr) {
    fkmc_chain_state& S = ctx->chain;
    const int N = ctx->N, C = S.n_chains;
    fkmc_prof_scope ps(ctx, "fu_ipr");
    fu_ipr_kernel<<<dim3((N + 255) / 256, C), 256, 0, ctx->stream>>>(N, C, S.fu_vt, S.fu_vslot, d_ipr);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}
