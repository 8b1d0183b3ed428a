// ORACLE (test infrastructure only) -- restatement of the dense symmetric eigensolver the
// reference calls at src/configuration.cpp:213 (Eigen::SelfAdjointEigenSolver, un-vendored and
// version-unpinned: cmake/CommonDefs.cmake:98 asks for "Eigen3 3.1").  Same algorithm shape:
// max-abs scaling, lower-triangle unblocked Householder tridiagonalisation (SYMV + SYR2 per
// column), implicit Wilkinson-shift QR on the tridiagonal, ascending selection sort.
// Also calc_ed's cache fill (src/configuration.cpp:226-244).
#include <cfloat>
#include <cmath>
#include <cstring>

#include "oracle.hpp"

namespace orc {

// Householder stage.  A: col-major n*n, lower triangle used and overwritten with the
// essential parts of the reflectors (below the sub-diagonal).  diag[n], sub[n-1].
void tridiagonalize_lower(int n, std::vector<double>& A, std::vector<double>& diag, std::vector<double>& sub,
                          std::vector<double>* hcoeffs) {
    diag.assign(n, 0.0);
    sub.assign(n > 1 ? n - 1 : 0, 0.0);
    if (hcoeffs) hcoeffs->assign(n > 1 ? n - 1 : 0, 0.0);
    std::vector<double> p(n), v(n);
    for (int i = 0; i < n - 1; ++i) {
        const int r = n - i - 1;
        double* x = &A[(size_t)i * n + i + 1];  // column i below the diagonal, length r
        double tail2 = 0;
        for (int k = 1; k < r; ++k) tail2 += x[k] * x[k];
        double c0 = x[0], tau, beta;
        if (tail2 <= DBL_MIN) {
            tau = 0;
            beta = c0;
            for (int k = 1; k < r; ++k) x[k] = 0;
        } else {
            beta = std::sqrt(c0 * c0 + tail2);
            if (c0 >= 0) beta = -beta;
            const double inv = 1.0 / (c0 - beta);
            for (int k = 1; k < r; ++k) x[k] *= inv;
            tau = (beta - c0) / beta;
        }
        v[0] = 1.0;
        for (int k = 1; k < r; ++k) v[k] = x[k];
        // p = tau * A22 * v, A22 = trailing r x r block, lower triangle only
        for (int k = 0; k < r; ++k) p[k] = 0;
        for (int c = 0; c < r; ++c) {
            const double* col = &A[(size_t)(i + 1 + c) * n + (i + 1)];  // col[k] = A22(k, c)
            const double vc = v[c];
            double dot = col[c] * vc;
            for (int k = c + 1; k < r; ++k) {
                p[k] += col[k] * vc;
                dot += col[k] * v[k];
            }
            p[c] += dot;
        }
        double pv = 0;
        for (int k = 0; k < r; ++k) {
            p[k] *= tau;
            pv += p[k] * v[k];
        }
        const double alpha = -0.5 * tau * pv;
        for (int k = 0; k < r; ++k) p[k] += alpha * v[k];
        // A22 -= v p^T + p v^T (lower)
        for (int c = 0; c < r; ++c) {
            double* col = &A[(size_t)(i + 1 + c) * n + (i + 1)];
            const double vc = v[c], pc = p[c];
            for (int k = c; k < r; ++k) col[k] -= v[k] * pc + p[k] * vc;
        }
        diag[i] = A[(size_t)i * n + i];
        sub[i] = beta;
        if (hcoeffs) (*hcoeffs)[i] = tau;
    }
    diag[n - 1] = A[(size_t)(n - 1) * n + (n - 1)];
}

static inline void givens(double p, double q, double& c, double& s) {
    if (q == 0) {
        c = p < 0 ? -1 : 1;
        s = 0;
    } else if (p == 0) {
        c = 0;
        s = q < 0 ? 1 : -1;
    } else if (std::fabs(p) > std::fabs(q)) {
        double t = q / p, u = std::sqrt(1 + t * t);
        if (p < 0) u = -u;
        c = 1 / u;
        s = -t * c;
    } else {
        double t = p / q, u = std::sqrt(1 + t * t);
        if (q < 0) u = -u;
        s = -1 / u;
        c = -t * s;
    }
}

// Implicit symmetric QR with Wilkinson shift on (diag, sub).  Q (optional) receives the
// rotations on its columns.  Returns 0 on convergence, 1 after 30*n sweeps.
int tridiag_ql_implicit(int n, double* diag, double* sub, double* Q) {
    int end = n - 1, start = 0, iter = 0;
    const double tiny = DBL_MIN, inv_eps = 1.0 / DBL_EPSILON;
    while (end > 0) {
        for (int i = start; i < end; ++i) {
            if (std::fabs(sub[i]) < tiny) {
                sub[i] = 0;
            } else {
                double sc = inv_eps * sub[i];
                if (sc * sc <= std::fabs(diag[i]) + std::fabs(diag[i + 1])) sub[i] = 0;
            }
        }
        while (end > 0 && sub[end - 1] == 0) end--;
        if (end <= 0) break;
        if (++iter > 30 * n) return 1;
        start = end - 1;
        while (start > 0 && sub[start - 1] != 0) start--;
        // one QR step on [start, end]
        double td = (diag[end - 1] - diag[end]) * 0.5, e = sub[end - 1], mu = diag[end];
        if (td == 0) {
            mu -= std::fabs(e);
        } else if (e != 0) {
            const double e2 = e * e, h = std::hypot(td, e);
            if (e2 == 0)
                mu -= e / ((td + (td > 0 ? h : -h)) / e);
            else
                mu -= e2 / (td + (td > 0 ? h : -h));
        }
        double x = diag[start] - mu, z = sub[start];
        for (int k = start; k < end && z != 0; ++k) {
            double c, s;
            givens(x, z, c, s);
            const double sdk = s * diag[k] + c * sub[k];
            const double dkp1 = s * sub[k] + c * diag[k + 1];
            diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
            diag[k + 1] = s * sdk + c * dkp1;
            sub[k] = c * sdk - s * dkp1;
            if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
            x = sub[k];
            if (k < end - 1) {
                z = -s * sub[k + 1];
                sub[k + 1] = c * sub[k + 1];
            }
            if (Q) {
                double* qk = Q + (size_t)k * n;
                double* qk1 = Q + (size_t)(k + 1) * n;
                for (int r = 0; r < n; ++r) {
                    const double a = qk[r], b = qk1[r];
                    qk[r] = c * a - s * b;
                    qk1[r] = s * a + c * b;
                }
            }
        }
    }
    return 0;
}

int eigh_lower(int n, const double* Ain, double* evals, double* evecs) {
    if (n == 1) {
        evals[0] = Ain[0];
        if (evecs) evecs[0] = 1;
        return 0;
    }
    std::vector<double> A(Ain, Ain + (size_t)n * n);
    double scale = 0;
    for (int c = 0; c < n; ++c)
        for (int r = c; r < n; ++r) scale = std::fmax(scale, std::fabs(A[(size_t)c * n + r]));
    if (scale == 0) scale = 1;
    for (int c = 0; c < n; ++c)
        for (int r = c; r < n; ++r) A[(size_t)c * n + r] /= scale;
    std::vector<double> diag, sub, hc;
    tridiagonalize_lower(n, A, diag, sub, &hc);
    std::vector<double> Q;
    if (evecs) {
        // Q = H_0 H_1 ... H_{n-2}, H_i = I - tau_i v_i v_i^T, v_i = [0..0, 1, A(i+2:, i)]
        Q.assign((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i) Q[(size_t)i * n + i] = 1;
        for (int i = n - 2; i >= 0; --i) {
            const int r = n - i - 1;
            const double* ess = &A[(size_t)i * n + i + 1];
            const double tau = hc[i];
            if (tau == 0) continue;
            for (int c = i + 1; c < n; ++c) {
                double* qc = &Q[(size_t)c * n + i + 1];
                double dot = qc[0];
                for (int k = 1; k < r; ++k) dot += ess[k] * qc[k];
                dot *= tau;
                qc[0] -= dot;
                for (int k = 1; k < r; ++k) qc[k] -= dot * ess[k];
            }
        }
    }
    int info = tridiag_ql_implicit(n, diag.data(), sub.data(), evecs ? Q.data() : nullptr);
    // ascending selection sort (columns follow)
    for (int i = 0; i < n - 1; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j)
            if (diag[j] < diag[k]) k = j;
        if (k != i) {
            std::swap(diag[i], diag[k]);
            if (evecs)
                for (int r = 0; r < n; ++r) std::swap(Q[(size_t)i * n + r], Q[(size_t)k * n + r]);
        }
    }
    for (int i = 0; i < n; ++i) evals[i] = diag[i] * scale;
    if (evecs) std::memcpy(evecs, Q.data(), sizeof(double) * (size_t)n * n);
    return info;
}

// src/configuration.cpp:226-244
double logz_from_spectrum(const std::vector<double>& spectrum, double beta, std::vector<double>* cexp,
                          std::vector<double>* cfermi) {
    const size_t n = spectrum.size();
    const double e0 = spectrum[0];
    const double logw0 = beta * e0;
    const double weight0 = std::exp(logw0);
    if (cexp) cexp->resize(n);
    if (cfermi) cfermi->resize(n);
    double logz = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const double e = spectrum[i];
        const double w = std::exp(-beta * (e - e0));
        const double ex = std::exp(beta * e);
        if (cexp) (*cexp)[i] = ex;
        if (cfermi) (*cfermi)[i] = 1.0 / (1.0 + ex);
        logz += std::log(weight0 + w) - logw0;
    }
    return logz;
}

// src/configuration.cpp:208-246
void calc_ed(const lattice& lat, const std::vector<int>& f, double U, double mu_c, double beta, bool evecs, ed_result& out) {
    std::vector<double> H;
    dense_hamiltonian(lat, f, U, mu_c, H);
    const int n = lat.N;
    out.spectrum.resize(n);
    if (evecs) out.evecs.resize((size_t)n * n); else out.evecs.clear();
    eigh_lower(n, H.data(), out.spectrum.data(), evecs ? out.evecs.data() : nullptr);
    out.logZ = logz_from_spectrum(out.spectrum, beta, &out.cached_exp, &out.cached_fermi);
}

}  // namespace orc
