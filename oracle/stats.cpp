// ORACLE (test infrastructure only) -- binning / jackknife restatement.
// Follows include/fk_mc/binning.hpp:89-171 and include/fk_mc/jackknife.hpp:50-82.
#include <cmath>
#include <functional>
#include <stdexcept>

#include "oracle.hpp"

namespace orc {

// binning.hpp:89-96
bin_stats calc_stats(const std::vector<double>& x) {
    const size_t n = x.size();
    double mean = 0;
    for (double v : x) mean += v;
    mean /= n;
    double var = 0;
    for (double v : x) var += (v - mean) * (v - mean);
    var /= (n - 1);
    return {double(n), mean, var, std::sqrt(var / n)};
}

// one level of binned_iterator (binning.hpp:26-44): pairwise averages, odd tail dropped
std::vector<double> bin_once(const std::vector<double>& x) {
    std::vector<double> out(x.size() / 2);
    for (size_t i = 0; i < out.size(); ++i) out[i] = (x[2 * i] + x[2 * i + 1]) / 2.;
    return out;
}

static std::vector<double> bin_depth(const std::vector<double>& x, int depth) {
    // the reference drops the tail with respect to the full step 2^depth (find_bin_range), not level by level
    const size_t step = size_t(1) << depth;
    if (step > x.size()) throw std::logic_error("Can't bin with binning step > container size");
    const size_t n = x.size() / step;
    std::vector<double> cur(x.begin(), x.begin() + n * step);
    for (int d = 0; d < depth; ++d) cur = bin_once(cur);
    return cur;
}

std::vector<bin_stats> accumulate_binning(const std::vector<double>& x, int max_depth) {
    std::vector<bin_stats> out;
    for (int d = 0; d <= max_depth; ++d) out.push_back(calc_stats(bin_depth(x, d)));
    return out;
}

// binning.hpp:163-171
double calc_cor_length(const std::vector<bin_stats>& b, int level) { return 0.5 * (std::pow(2., level) * b[level].var / b[0].var - 1); }

}  // namespace orc

extern "C" {

int orc_binning(int n, const double* x, int max_depth, double* rows /*[(max_depth+1)][5]: n, mean, var, err, tau*/) {
    try {
        std::vector<double> v(x, x + n);
        auto r = orc::accumulate_binning(v, max_depth);
        for (int d = 0; d <= max_depth; ++d) {
            rows[5 * d + 0] = r[d].n; rows[5 * d + 1] = r[d].mean; rows[5 * d + 2] = r[d].var; rows[5 * d + 3] = r[d].err;
            rows[5 * d + 4] = orc::calc_cor_length(r, d);
        }
    } catch (std::exception&) { return -1; }
    return 0;
}

// jackknife of f(e, e2, de2) = e2 - de2 - e*e (the reference's cv functor without the beta^2/N prefactor,
// test/jackknife_test.cpp:95) or of f(x) = x (nseries == 1) at one bin depth -- jackknife.hpp:50-82
int orc_jackknife(int n, int nseries, const double* data /*[nseries][n]*/, int depth, double* out4) {
    try {
        std::vector<std::vector<double>> d(nseries);
        for (int s = 0; s < nseries; ++s) d[s] = orc::bin_depth(std::vector<double>(data + (size_t)s * n, data + (size_t)(s + 1) * n), depth);
        auto F = [nseries](const std::vector<double>& a) { return nseries == 1 ? a[0] : a[1] - a[2] - a[0] * a[0]; };
        const size_t size = d[0].size();
        std::vector<double> means(nseries);
        for (int s = 0; s < nseries; ++s) means[s] = orc::calc_stats(d[s]).mean;
        const double U0 = F(means);
        std::vector<double> U(size), loo(nseries);
        for (size_t j = 0; j < size; ++j) {
            for (int s = 0; s < nseries; ++s) loo[s] = (double(size) * means[s] - d[s][j]) / double(size - 1);
            U[j] = F(loo);
        }
        auto st = orc::calc_stats(U);
        const double Uavg = U0 - (size - 1) * (st.mean - U0);
        const double dU = (size - 1) * st.err;
        out4[0] = double(size); out4[1] = Uavg; out4[2] = dU * dU * size; out4[3] = dU;
    } catch (std::exception&) { return -1; }
    return 0;
}
}
