// ORACLE (test infrastructure only) -- Metropolis driver, moves and measures.
// Follows src/mc_metropolis.cpp:34-61 (update/measure), src/moves.cpp, src/moves_chebyshev.cpp,
// include/fk_mc/fk_mc.hxx:35-125 (initialize: move/measure registration order),
// src/measures/energy.cpp:6-26, include/fk_mc/measures/ipr.hpp:39-56.
// libstdc++ <random> is used directly so RNG consumption is the reference's by construction.
#include <cmath>
#include <functional>
#include <memory>
#include <numeric>
#include <stdexcept>

#include "oracle.hpp"

namespace orc {

// src/measures/energy.cpp:6-26
void measure_energy(const ed_result& ed, double beta, double mu_f, int nf, double eff, double& e, double& d2e, double& ec) {
    (void)beta;
    const size_t n = ed.spectrum.size();
    double s1 = 0, s2 = 0;
    for (size_t i = 0; i < n; ++i) {
        const double ex = ed.cached_exp[i];
        s1 += ed.spectrum[i] / (1.0 + ex);
        s2 += ed.spectrum[i] * ed.spectrum[i] / (1.0 + 0.5 * (ex + 1. / ex));
    }
    ec = s1;
    e = s1 - mu_f * nf + eff;
    d2e = s2 / 2.0;
}

// include/fk_mc/measures/ipr.hpp:39-56: ||psi||_4 / ||psi||_2^2 per eigenvector (column)
void measure_ipr(int n, const std::vector<double>& evecs, std::vector<double>& ipr) {
    ipr.resize(n);
    for (int k = 0; k < n; ++k) {
        double s2 = 0, s4 = 0;
        for (int r = 0; r < n; ++r) {
            double x = evecs[(size_t)k * n + r];
            s2 += x * x;
            s4 += x * x * x * x;
        }
        ipr[k] = std::pow(s4, 0.25) / s2;
    }
}

namespace {

struct config_state {
    std::vector<int> f;
    bool ed_valid = false, cheb_valid = false;
    ed_result ed;
    cheb_result cheb;
    int nf() const { return std::accumulate(f.begin(), f.end(), 0); }
};

struct engine {
    const mc_params& p;
    lattice lat;
    std::unique_ptr<chebyshev_eval> cheb;
    random_generator random;
    config_state config, new_config;
    std::vector<double> W;  // f-f interaction (1-D only)

    engine(const mc_params& p_, int rank) : p(p_), lat(make_lattice(p_.kind, p_.L, p_.t, p_.tp)), random(p_.seed + rank) {
        if (lat.ndim == 1) W.assign(p_.W, p_.W + p_.n_W);  // fk_mc.hxx:40-45: W only reaches the configuration on 1-D lattices
    }

    void ensure_ed(config_state& c, bool evecs = false) {
        if (c.ed_valid && (!evecs || !c.ed.evecs.empty())) return;
        calc_ed(lat, c.f, p.U, p.mu_c, p.beta, evecs, c.ed);
        c.ed_valid = true;
    }
    void ensure_cheb(config_state& c) {
        if (c.cheb_valid) return;
        calc_chebyshev(lat, c.f, p.U, p.mu_c, p.beta, *cheb, p.emode, false, c.cheb);
        c.cheb_valid = true;
    }
    void reset_cache(config_state& c) { c.ed_valid = false; c.cheb_valid = false; c.ed.evecs.clear(); }
    double logz(config_state& c) {
        if (p.cheb_moves) { ensure_cheb(c); return c.cheb.logZ; }
        ensure_ed(c); return c.ed.logZ;
    }
};

}  // namespace

void mc_run(const mc_params& p, int rank, mc_result& res, mc_trace* trace) {
    engine E(p, rank);
    const int V = E.lat.N;
    const double beta = p.beta;
    // fk_mc.hxx:46,53
    randomize_f(E.random, V, (size_t)p.nf_start, E.config.f);
    E.reset_cache(E.config);
    if (p.cheb_moves) {
        int M, G;
        cheb_sizes(V, p.cheb_prefactor, M, G);
        E.cheb.reset(new chebyshev_eval(M, G));
    }
    // move registry (fk_mc.hxx:67-78): flip, add_remove, reshuffle in this order, each only if weight > eps
    enum { FLIP = 0, ADDREM = 1, RESHUFFLE = 2 };
    std::vector<int> moves;
    std::vector<double> probs;
    const double eps = std::numeric_limits<double>::epsilon();
    if (p.mc_flip > eps) { moves.push_back(FLIP); probs.push_back(p.mc_flip); }
    if (p.mc_add_remove > eps) { moves.push_back(ADDREM); probs.push_back(p.mc_add_remove); }
    if (p.mc_reshuffle > eps) { moves.push_back(RESHUFFLE); probs.push_back(p.mc_reshuffle); }
    if (moves.empty()) throw std::logic_error("No registered moves");
    std::discrete_distribution<> move_distrib(probs.begin(), probs.end());
    std::uniform_real_distribution<> metropolis_distrib(0, 1);
    const double exp_beta_mu_f = std::exp(beta * p.mu_f);  // moves.hpp:49

    const long total_sweeps = (long)p.nsweeps + p.ntherm_sweeps;
    long measure_count = 0;
    res = mc_result();
    res.spectrum_avg.assign(V, 0.0);
    int specZ = 0;
    if (trace) *trace = mc_trace();

    for (long sweep = 0; sweep < total_sweeps; ++sweep) {
        // ---- update(): src/mc_metropolis.cpp:34-52 ----
        for (int m = 0; m < p.sweep_len; ++m) {
            const int move_index = move_distrib(E.random);
            const int kind = moves[move_index];
            double weight = 0;
            int sa = -1, sb = -1;
            config_state& config = E.config;
            config_state& nc = E.new_config;
            std::uniform_int_distribution<> distr(0, V - 1);
            if (kind == FLIP) {
                // src/moves.cpp:5-21 / src/moves_chebyshev.cpp:6-24
                if (p.cheb_moves) E.ensure_cheb(config);
                const int nf = config.nf();
                if (nf == 0 || nf == V) {
                    weight = 0;
                } else {
                    nc = config;
                    size_t from = distr(E.random); while (nc.f[from] == 0) from = distr(E.random);
                    size_t to = distr(E.random); while (nc.f[to] == 1) to = distr(E.random);
                    sa = (int)from; sb = (int)to;
                    const double lz_old = E.logz(config);
                    nc.f[from] = 0; nc.f[to] = 1;
                    E.reset_cache(nc);
                    const double lz_new = E.logz(nc);
                    const double ff_diff = calc_ff_energy(E.lat.ndim, nc.f, E.W) - calc_ff_energy(E.lat.ndim, config.f, E.W);
                    weight = std::exp(lz_new - lz_old - beta * ff_diff);  // Q7: E_ff included uniformly (0 for D>=2)
                }
            } else if (kind == ADDREM) {
                // src/moves.cpp:52-67 / src/moves_chebyshev.cpp:55-72
                const double lz_old0 = p.cheb_moves ? E.logz(config) : 0;  // cheb variant evaluates the cache first
                (void)lz_old0;
                nc = config;
                size_t to = distr(E.random);
                sa = (int)to;
                nc.f[to] = 1 - config.f[to];
                const double lz_old = E.logz(config);
                E.reset_cache(nc);
                const double lz_new = E.logz(nc);
                const double ff_diff = calc_ff_energy(E.lat.ndim, nc.f, E.W) - calc_ff_energy(E.lat.ndim, config.f, E.W);
                const double ratio = std::exp(lz_new - lz_old);
                weight = (nc.f[to] ? ratio * exp_beta_mu_f : ratio / exp_beta_mu_f) * std::exp(-beta * ff_diff);
            } else {
                // src/moves.cpp:35-49 / src/moves_chebyshev.cpp:38-52 with signed nf difference (Q3)
                if (p.cheb_moves) E.ensure_cheb(config);
                nc = config;
                randomize_f(E.random, V, 0, nc.f);
                E.reset_cache(nc);
                const double lz_old = E.logz(config);
                const double lz_new = E.logz(nc);
                const double log_ratio = lz_new - lz_old;
                const double ff_diff = calc_ff_energy(E.lat.ndim, nc.f, E.W) - calc_ff_energy(E.lat.ndim, config.f, E.W);
                const double dn = double(nc.nf()) - double(config.nf());
                if (beta * p.mu_f * dn - ff_diff > 2.7182818 - log_ratio) weight = 1;
                else if (beta * p.mu_f * dn - ff_diff + log_ratio < 0) weight = 0;
                else weight = std::exp(log_ratio) * std::exp(beta * (p.mu_f * dn - ff_diff));
            }
            const double u = metropolis_distrib(E.random);
            const bool acc = std::fabs(weight) > u;
            if (trace) {
                trace->move.push_back(kind);
                trace->site_a.push_back(sa);
                trace->site_b.push_back(sb);
                trace->weight.push_back(weight);
                trace->u.push_back(u);
                trace->accepted.push_back(acc ? 1 : 0);
                trace->logz_new.push_back((sa >= 0 || kind == RESHUFFLE) ? (p.cheb_moves ? nc.cheb.logZ : nc.ed.logZ) : 0.0);
            }
            if (acc) {
                E.config = nc;  // moves.cpp:23-28
                res.naccept++;
            }
        }
        // ---- measure(): src/mc_metropolis.cpp:54-61 ----
        if (measure_count >= p.ntherm_sweeps) {
            config_state& config = E.config;
            if (p.measure_ipr) {
                E.ensure_ed(config, true);
                std::vector<double> ipr;
                measure_ipr(V, config.ed.evecs, ipr);
                res.ipr_history.push_back(ipr);
            }
            if (p.measure_energy) {
                E.ensure_ed(config);
                double e, d2e, ec;
                measure_energy(config.ed, beta, p.mu_f, config.nf(), calc_ff_energy(E.lat.ndim, config.f, E.W), e, d2e, ec);
                res.energies.push_back(e);
                res.d2energies.push_back(d2e);
                res.c_energies.push_back(ec);
                // src/measures/spectrum.cpp:13-21
                for (int i = 0; i < V; ++i) res.spectrum_avg[i] = (res.spectrum_avg[i] * specZ + config.ed.spectrum[i]) / (specZ + 1);
                specZ++;
                res.spectrum_history.push_back(config.ed.spectrum);
            }
            res.focc_history.push_back(config.f);
            res.nf_series.push_back(config.nf());
        }
        measure_count++;
    }
    res.f_final = E.config.f;
    res.logz_final = p.cheb_moves ? (E.config.cheb_valid ? E.config.cheb.logZ : 0) : (E.config.ed_valid ? E.config.ed.logZ : 0);
}

}  // namespace orc

namespace orc {

// include/fk_mc/measures/stiffness.hpp:69-187 (2-D hypercubic, t_ = 1): T = -pi sum_k (V^T Tm V)_kk f_k,
// V = sum_{i>j, |e_i-e_j|>1e-12, |sigma|>1e-13} 2 pi (f_j - f_i) mJ_ji mJ_ij / (e_i - e_j), stiffness = (T + V) / N;
// conductivity sigma(w) = sum of Lorentzians of the resonant terms.
double measure_stiffness(const lattice& lat, const ed_result& ed, double beta, double offset, const std::vector<double>* wgrid,
                         std::vector<double>* cond) {
    (void)beta;
    const int n = lat.N, L = lat.L;
    if (lat.ndim < 2) throw std::logic_error("no stiffness in 1-D");
    const std::vector<double>& V = ed.evecs;
    // left / right neighbours along dimension 0 (the first coordinate)
    std::vector<int> left(n), right(n);
    for (int i = 0; i < n; ++i) {
        auto pos = lat.index_to_pos(i);
        auto pl = pos, pr = pos;
        pl[0] = ((pos[0] - 1) + L) % L;
        pr[0] = ((pos[0] + 1) + L) % L;
        left[i] = lat.pos_to_index(pl);
        right[i] = lat.pos_to_index(pr);
    }
    // TV = Tm V, JV = Jm V with Tm(i,left)=Tm(i,right)=-1, Jm(i,left)=-1, Jm(i,right)=+1
    std::vector<double> TV((size_t)n * n), JV((size_t)n * n);
    for (int k = 0; k < n; ++k)
        for (int i = 0; i < n; ++i) {
            const double vl = V[(size_t)k * n + left[i]], vr = V[(size_t)k * n + right[i]];
            TV[(size_t)k * n + i] = -vl - vr;
            JV[(size_t)k * n + i] = -vl + vr;
        }
    double T = 0;
    for (int k = 0; k < n; ++k) {
        double d = 0;
        for (int i = 0; i < n; ++i) d += V[(size_t)k * n + i] * TV[(size_t)k * n + i];
        T += d * ed.cached_fermi[k];
    }
    T *= -M_PI;
    double Vsum = 0;
    std::vector<std::pair<double, double>> terms;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            if (!(std::fabs(ed.spectrum[i] - ed.spectrum[j]) > 1e-12)) continue;
            double mji = 0, mij = 0;  // mJ(j,i) = v_j . (Jm v_i)
            for (int r = 0; r < n; ++r) {
                mji += V[(size_t)j * n + r] * JV[(size_t)i * n + r];
                mij += V[(size_t)i * n + r] * JV[(size_t)j * n + r];
            }
            const double sigma_v = M_PI * (ed.cached_fermi[j] - ed.cached_fermi[i]) * mji * mij;
            if (std::fabs(sigma_v) > 1e-13) {
                Vsum += 2. * sigma_v / (ed.spectrum[i] - ed.spectrum[j]);
                terms.push_back({ed.spectrum[j] - ed.spectrum[i], sigma_v});
                terms.push_back({ed.spectrum[i] - ed.spectrum[j], -sigma_v});
            }
        }
    if (wgrid && cond) {
        cond->assign(wgrid->size(), 0.0);
        for (size_t w = 0; w < wgrid->size(); ++w)
            for (auto& t : terms) {
                const double x = (*wgrid)[w] + t.first;
                (*cond)[w] += offset / M_PI / (x * x + offset * offset) * t.second;
            }
    }
    return (Vsum + T) / n;
}

}  // namespace orc

extern "C" int orc_stiffness(int kind, int L, double t, const int* f, double U, double mu_c, double beta, double offset, int nw,
                             const double* wgrid, double* cond, double* stiffness) {
    try {
        orc::lattice l = orc::make_lattice(kind, L, t, 1.0);
        orc::ed_result ed;
        orc::calc_ed(l, std::vector<int>(f, f + l.N), U, mu_c, beta, true, ed);
        std::vector<double> wg(wgrid, wgrid + nw), cd;
        *stiffness = orc::measure_stiffness(l, ed, beta, offset, &wg, &cd);
        for (int i = 0; i < nw; ++i) cond[i] = cd[i];
    } catch (std::exception&) { return -1; }
    return 0;
}
