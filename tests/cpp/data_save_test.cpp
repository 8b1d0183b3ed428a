// CPU-only check of include/fk_mc_b200/data_save.hpp: writes an output file with the reference's layout from a deterministic
// series; tests/test_h5out.py reads it back with the Python reader and compares with fk_mc_b200/stats.py.
//   data_save_test <out.h5> <n> <beta> <volume>
#include <cstdio>
#include <cstdlib>

#include "fk_mc_b200/data_save.hpp"

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    const size_t n = std::strtoul(argv[2], nullptr, 10);
    const double beta = std::atof(argv[3]), volume = std::atof(argv[4]);
    // correlated pseudo-random series (AR(1) on a 64-bit LCG), reproduced by the Python side
    std::vector<double> e(n), d2(n), ce(n);
    uint64_t s = 88172645463325252ull;
    double x = 0.0;
    for (size_t i = 0; i < n; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const double u = double(s >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        x = 0.7 * x + u;
        e[i] = -0.25 + 0.01 * x;
        d2[i] = 0.02 + 0.001 * u;
        ce[i] = e[i] + 0.5;
    }
    fk::param_map p;
    p["beta"] = beta;
    p["U"] = 1.0;
    p["L"] = int64_t(8);
    p["nsweeps"] = int64_t(n);
    p["output"] = std::string("output.h5");
    std::vector<double> hist(3 * 4);
    for (size_t i = 0; i < hist.size(); ++i) hist[i] = double(i) * 0.5;
    const auto out = fk::save_all_data(argv[1], p, e, d2, ce, beta, volume, -1, {{"ipr_history", {hist, {3, 4}}}});
    for (auto& [name, st] : out) std::printf("%s %.17g %.17g %.17g %.17g %zu\n", name.c_str(), st.stats[0], st.stats[1], st.stats[2], st.stats[3], st.binning.size());
    return 0;
}
