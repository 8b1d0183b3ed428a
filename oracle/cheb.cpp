// ORACLE (test infrastructure only) -- Chebyshev/KPM weight evaluation.
// Follows include/fk_mc/chebyshev.hpp (evaluator) and src/configuration.cpp:94-205
// (calc_chebyshev).  The two ARPACK calls at src/configuration.cpp:99-100 are replaced either by
// the extremes of the dense spectrum (emode 0) or by a Lanczos iteration with full
// re-orthogonalisation (emode 1, the CPU-baseline stand-in for ARPACK nev=1 SA/LA, tol=0).
#include <cfloat>
#include <cmath>
#include <stdexcept>

#include "oracle.hpp"

namespace orc {

// include/fk_mc/chebyshev.hpp:21-34
chebyshev_eval::chebyshev_eval(int max_moment, int grid_size)
    : M(max_moment + max_moment % 2), G(grid_size), angle_grid(grid_size), lobatto_grid(grid_size),
      chebt((size_t)(max_moment + max_moment % 2) * grid_size) {
    for (int i = 0; i < G; ++i) {
        angle_grid[i] = (i == G - 1) ? 1.0 : double(i) * (1.0 / double(G - 1));
        lobatto_grid[i] = -std::cos(M_PI * angle_grid[i]);
        for (int k = 0; k < M; ++k) chebt[(size_t)k * G + i] = std::cos(k * std::acos(lobatto_grid[i]));
    }
}

// include/fk_mc/chebyshev.hpp:45-54 (trapezoid rule in the angle variable)
double chebyshev_eval::moment(const std::vector<double>& in, int order) const {
    double s = 0.0;
    const double* T = &chebt[(size_t)order * G];
    for (int i = 0; i < G - 1; ++i) s += (in[i + 1] * T[i + 1] + in[i] * T[i]) * (angle_grid[i + 1] - angle_grid[i]);
    return s * 0.5;
}

// include/fk_mc/fk_mc.hxx:60-63
void cheb_sizes(int msize, double prefactor, int& M, int& G) {
    int cheb_size = int(std::log(double(msize)) * prefactor);
    cheb_size += cheb_size % 2;
    M = cheb_size;
    G = std::max(cheb_size * 2, 10);
}

static inline void spmv(const lattice& lat, const std::vector<double>& diag, const double* x, double* y) {
    const int n = lat.N;
    for (int i = 0; i < n; ++i) {
        double s = diag[i] * x[i];
        for (auto& e : lat.rows[i]) s += e.second * x[e.first];
        y[i] = s;
    }
}

// number of eigenvalues of the symmetric tridiagonal (d, e) that are < x
static int sturm_count(int k, const double* d, const double* e, double x, double pivmin) {
    int cnt = 0;
    double q = d[0] - x;
    if (std::fabs(q) < pivmin) q = -pivmin;
    if (q < 0) cnt++;
    for (int i = 1; i < k; ++i) {
        q = d[i] - x - e[i - 1] * e[i - 1] / q;
        if (std::fabs(q) < pivmin) q = -pivmin;
        if (q < 0) cnt++;
    }
    return cnt;
}

// idx-th (0-based, ascending) eigenvalue of the tridiagonal by bisection
static double tridiag_kth(int k, const double* d, const double* e, int idx) {
    double lo = d[0], hi = d[0], emax = 0;
    for (int i = 0; i < k; ++i) {
        double r = (i > 0 ? std::fabs(e[i - 1]) : 0) + (i < k - 1 ? std::fabs(e[i]) : 0);
        lo = std::fmin(lo, d[i] - r);
        hi = std::fmax(hi, d[i] + r);
        if (i < k - 1) emax = std::fmax(emax, e[i] * e[i]);
    }
    const double pivmin = std::fmax(DBL_MIN * std::fmax(1.0, emax), DBL_MIN);
    const double span = hi - lo;
    lo -= 2 * DBL_EPSILON * span + 2 * pivmin;
    hi += 2 * DBL_EPSILON * span + 2 * pivmin;
    for (int it = 0; it < 200; ++it) {
        double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        if (sturm_count(k, d, e, mid, pivmin) > idx) hi = mid; else lo = mid;
    }
    return 0.5 * (lo + hi);
}

void lanczos_extremal(const lattice& lat, const std::vector<double>& diag, double& e_min, double& e_max, int& steps) {
    const int n = lat.N;
    const int kmax = n;
    std::vector<std::vector<double>> V;
    std::vector<double> alpha, beta, w(n);
    std::mt19937 g(20240229u);
    std::vector<double> v(n);
    double nrm = 0;
    for (int i = 0; i < n; ++i) {
        v[i] = double(g()) / 4294967296.0 - 0.5;
        nrm += v[i] * v[i];
    }
    nrm = std::sqrt(nrm);
    for (auto& x : v) x /= nrm;
    V.push_back(v);
    double prev_lo = 0, prev_hi = 0;
    int stagn = 0;
    double scale = 0;
    for (int k = 0; k < kmax; ++k) {
        spmv(lat, diag, V[k].data(), w.data());
        double a = 0;
        for (int i = 0; i < n; ++i) a += V[k][i] * w[i];
        alpha.push_back(a);
        for (int i = 0; i < n; ++i) w[i] -= a * V[k][i];
        if (k > 0)
            for (int i = 0; i < n; ++i) w[i] -= beta[k - 1] * V[k - 1][i];
        for (int pass = 0; pass < 2; ++pass)
            for (int j = 0; j <= k; ++j) {
                double dot = 0;
                for (int i = 0; i < n; ++i) dot += V[j][i] * w[i];
                for (int i = 0; i < n; ++i) w[i] -= dot * V[j][i];
            }
        double b = 0;
        for (int i = 0; i < n; ++i) b += w[i] * w[i];
        b = std::sqrt(b);
        scale = std::fmax(scale, std::fabs(a) + b);
        const int kk = k + 1;
        bool last = (kk == kmax) || (b <= 1e-13 * scale);
        if (last || (kk >= 8 && kk % 4 == 0)) {
            double lo = tridiag_kth(kk, alpha.data(), beta.data(), 0);
            double hi = tridiag_kth(kk, alpha.data(), beta.data(), kk - 1);
            if (kk > 8 && std::fabs(lo - prev_lo) <= 4 * DBL_EPSILON * scale && std::fabs(hi - prev_hi) <= 4 * DBL_EPSILON * scale)
                stagn++;
            else
                stagn = 0;
            prev_lo = lo;
            prev_hi = hi;
            if (last || stagn >= 2) {
                e_min = lo;
                e_max = hi;
                steps = kk;
                return;
            }
        }
        beta.push_back(b);
        for (int i = 0; i < n; ++i) v[i] = w[i] / b;
        V.push_back(v);
    }
    throw std::logic_error("oracle: lanczos did not terminate");
}

// src/configuration.cpp:94-205
void calc_chebyshev(const lattice& lat, const std::vector<int>& f, double U, double mu_c, double beta,
                    const chebyshev_eval& cheb, int emode, bool prune, cheb_result& out) {
    if (lat.kind == HONEYCOMB_REF) throw std::logic_error("oracle: KPM on the non-symmetric literal honeycomb is not meaningful (Q1)");
    const int n = lat.N;
    std::vector<double> diag(n);
    for (int i = 0; i < n; ++i) diag[i] = -mu_c + U * f[i] + lat.hop(i, i);
    double e_min, e_max;
    if (emode == 0) {
        ed_result ed;
        calc_ed(lat, f, U, mu_c, beta, false, ed);
        e_min = ed.spectrum.front();
        e_max = ed.spectrum.back();
    } else {
        lanczos_extremal(lat, diag, e_min, e_max, out.lanczos_steps);
    }
    const double a = (e_max - e_min) / 2., b = (e_max + e_min) / 2.;
    out.e_min = e_min; out.e_max = e_max; out.a = a; out.b = b;

    const int M = cheb.M;
    if (M % 2) throw std::logic_error("cheb_size must be even");
    // x = (H - b)/a as (xdiag, scaled neighbour lists)
    std::vector<double> xdiag(n);
    double trx = 0;
    for (int i = 0; i < n; ++i) {
        xdiag[i] = (diag[i] - b) / a;
        trx += xdiag[i];
    }
    auto apply_x = [&](const double* v, double* y) {
        for (int i = 0; i < n; ++i) {
            double s = xdiag[i] * v[i];
            for (auto& e : lat.rows[i])
                if (e.first != i) s += (e.second / a) * v[e.first];
            y[i] = s;
        }
    };
    const int half = M / 2;
    std::vector<double> trT(half + 1, 0.0), dot01(half + 1, 0.0), dot11(half + 1, 0.0);
    std::vector<bool> sparse_level(half + 1, true);
    if (!prune) {
        std::vector<double> v0(n), v1(n), v2(n);
        for (int j = 0; j < n; ++j) {
            for (int i = 0; i < n; ++i) v0[i] = 0;
            v0[j] = 1;
            apply_x(v0.data(), v1.data());
            for (int m = 2; m <= half; ++m) {
                apply_x(v1.data(), v2.data());
                double d01 = 0, d11 = 0;
                for (int i = 0; i < n; ++i) {
                    v2[i] = 2. * v2[i] - v0[i];
                    d01 += v1[i] * v2[i];
                    d11 += v2[i] * v2[i];
                }
                trT[m] += v2[j];
                dot01[m] += d01;
                dot11[m] += d11;
                v0.swap(v1);
                v1.swap(v2);
            }
        }
    } else {
        // level-by-level with Eigen's pruned(1.0) (drops |x| <= 1e-12) while the iterate is < 50 % full
        std::vector<double> T0((size_t)n * n, 0.0), T1((size_t)n * n, 0.0), T2((size_t)n * n);
        for (int j = 0; j < n; ++j) {
            T0[(size_t)j * n + j] = 1;
            apply_x(&T0[(size_t)j * n], &T1[(size_t)j * n]);
            for (int i = 0; i < n; ++i)
                if (std::fabs(T1[(size_t)j * n + i]) <= 1e-12) T1[(size_t)j * n + i] = 0;
        }
        bool still_sparse = true;
        for (int m = 2; m <= half; ++m) {
            size_t nnz = 0;
            for (int j = 0; j < n; ++j) {
                double* c2 = &T2[(size_t)j * n];
                apply_x(&T1[(size_t)j * n], c2);
                double d01 = 0, d11 = 0;
                for (int i = 0; i < n; ++i) {
                    double t = 2. * c2[i];
                    if (still_sparse && std::fabs(t) <= 1e-12) t = 0;
                    c2[i] = t - T0[(size_t)j * n + i];
                    if (c2[i] != 0) nnz++;
                    d01 += T1[(size_t)j * n + i] * c2[i];
                    d11 += c2[i] * c2[i];
                }
                trT[m] += c2[j];
                dot01[m] += d01;
                dot11[m] += d11;
            }
            T0.swap(T1);
            T1.swap(T2);
            if (still_sparse) still_sparse = (double(nnz) / n / n < 0.5);
        }
    }
    out.moments.assign(M, 0.0);
    std::vector<bool> is_set(M, false);
    out.moments[0] = 1.0; is_set[0] = true;
    out.moments[1] = trx / n; is_set[1] = true;
    for (int m = 2; m <= half; ++m) {
        if (!is_set[m]) {
            out.moments[m] = trT[m] / n;
            is_set[m] = true;
        }
        int k = 2 * m - 1;
        if (k < M && k >= half) {
            out.moments[k] = (dot01[m] * 2. - trx) / n;
            is_set[k] = true;
            if (k != M - 1) {
                ++k;
                out.moments[k] = (dot11[m] / n * 2. - 1.0);
                is_set[k] = true;
            }
        }
    }
    auto logz_f = [a, b, beta, n](double w) { return n * std::log(1. + std::exp(-beta * (a * w + b))); };
    double s = cheb.moment_f(logz_f, 0);
    for (int m = 1; m < M; ++m) s += 2. * cheb.moment_f(logz_f, m) * out.moments[m];
    out.logZ = s;
}

}  // namespace orc
