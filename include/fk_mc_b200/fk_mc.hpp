// fk_mc_b200/fk_mc.hpp -- C++ host mirror of the fk_mc interfaces on the weight-evaluation path, on top of
// the C ABI (fkmc.h).  Same names, argument meaning and error behaviour as the reference:
//   fk::hypercubic_lattice<D>, fill_nearest_neighbors / fill_triangular / fill_honeycomb   (include/fk_mc/lattice/hypercubic.hpp)
//   fk::config_params, ed_cache, chebyshev_cache, configuration_t                          (include/fk_mc/configuration.hpp)
//   fk::chebyshev::chebyshev_eval                                                          (include/fk_mc/chebyshev.hpp)
//   fk::move_flip / move_addremove / move_randomize and fk::chebyshev::move_*              (include/fk_mc/moves*.hpp)
//   fk::measure_energy                                                                     (include/fk_mc/measures/energy.hpp)
//   alps::mc_metropolis (add_move / add_measure / update / measure / rng)                  (include/fk_mc/mc_metropolis.hpp)
// Eigen / ALPSCore / Boost.MPI are not needed: arrays are std::vector, errors are the same exception types.
// Everything numerical happens on the GPU behind fkmc_logz_ed_batched / fkmc_logz_kpm_batched (batch of one here; the
// batched product is fkmc_chain_* / fk::batched_chains below).  Header-only; link with -lfkmc_b200.
#pragma once
#include <array>
#include <cmath>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../fkmc.h"

namespace fk {

typedef std::mt19937 random_generator;  // include/fk_mc/common.hpp:18

// status code -> the reference's exception types (FKMC_ERROR throws std::logic_error, common.hpp:50-59)
inline void fkmc_check(int rc, const fkmc_ctx* ctx) {
    if (rc == FKMC_OK) return;
    const std::string msg = fkmc_last_error(ctx);
    if (rc == FKMC_ERR_INVALID || rc == FKMC_ERR_STATE) throw std::logic_error(msg);
    throw std::runtime_error(msg);
}

/// Lattice of a finite volume (abstract_lattice, include/fk_mc/lattice.hpp:12-46) backed by an fkmc context.
struct abstract_lattice {
    virtual ~abstract_lattice() { if (ctx_) fkmc_destroy(ctx_); }
    size_t msize() const { return size_t(fkmc_volume(ctx())); }
    size_t norbs() const { return 1; }
    size_t volume() const { return msize(); }
    virtual size_t ndim() const = 0;
    /// dense row-major hopping matrix (the reference returns an Eigen::SparseMatrix)
    std::vector<double> hopping_m() const {
        std::vector<double> H(msize() * msize());
        fkmc_check(fkmc_hopping_dense(ctx(), H.data()), ctx());
        return H;
    }
    fkmc_ctx* ctx() const {
        if (!ctx_) throw std::logic_error("Failed to initialize lattice. ");  // lattice.hpp:52
        return ctx_;
    }
    void build(int kind, int L, double t, double tp, int device = 0, int max_batch = 1) {
        if (ctx_) fkmc_destroy(ctx_);
        ctx_ = nullptr;
        int rc = fkmc_create(&ctx_, device, kind, L, t, tp, max_batch);
        if (rc == FKMC_ERR_INVALID) throw std::logic_error(fkmc_last_error(nullptr));
        if (rc) throw std::runtime_error(fkmc_last_error(nullptr));
    }
    abstract_lattice() = default;
    abstract_lattice(const abstract_lattice&) = delete;
    abstract_lattice& operator=(const abstract_lattice&) = delete;

protected:
    fkmc_ctx* ctx_ = nullptr;
};

template <size_t D>
class hypercubic_lattice : public abstract_lattice {
public:
    static constexpr size_t Ndim = D;
    typedef std::array<int, D> pos_t;
    explicit hypercubic_lattice(size_t lattice_size) { dims_.fill(int(lattice_size)); }
    size_t ndim() const override { return D; }
    std::array<int, D> const& dims() const { return dims_; }
    /// src/lattice/hypercubic.cpp:31-39 (last coordinate fastest)
    pos_t index_to_pos(size_t index) const {
        pos_t out;
        for (int i = int(D) - 1; i >= 0; i--) { out[i] = int(index % dims_[i]); index /= dims_[i]; }
        return out;
    }
    /// src/lattice/hypercubic.cpp:42-51
    size_t pos_to_index(pos_t pos) const {
        size_t out = 0, mult = 1;
        for (int i = int(D) - 1; i >= 0; i--) { out += pos[i] * mult; mult *= dims_[i]; }
        return out;
    }

protected:
    std::array<int, D> dims_;
};

/// src/lattice/hypercubic.cpp:116-131
template <size_t D>
hypercubic_lattice<D>& fill_nearest_neighbors(hypercubic_lattice<D>& l, double t, int device = 0, int max_batch = 1) {
    l.build(D == 1 ? FKMC_CUBIC1D : (D == 2 ? FKMC_CUBIC2D : FKMC_CUBIC3D), l.dims()[0], t, 0.0, device, max_batch);
    return l;
}
/// src/lattice/hypercubic.cpp:137-155
inline hypercubic_lattice<2>& fill_triangular(hypercubic_lattice<2>& l, double t, double tp, int device = 0, int max_batch = 1) {
    l.build(FKMC_TRIANGULAR, l.dims()[0], t, tp, device, max_batch);
    return l;
}
/// src/lattice/hypercubic.cpp:160-203 (intended symmetric brick wall; literal_lower = what Eigen's solver sees of the literal matrix)
inline hypercubic_lattice<2>& fill_honeycomb(hypercubic_lattice<2>& l, double t, bool literal_lower = false, int device = 0, int max_batch = 1) {
    if (l.dims()[0] % 2 != 0) throw std::logic_error("Need even size");
    l.build(literal_lower ? FKMC_HONEYCOMB_REF_LOWER : FKMC_HONEYCOMB, l.dims()[0], t, 0.0, device, max_batch);
    return l;
}

namespace chebyshev {
/// include/fk_mc/chebyshev.hpp:15-63 (host copy of the evaluator; the device keeps its own tables)
struct chebyshev_eval {
    int cheb_size() const { return M_; }
    int grid_size() const { return G_; }
    chebyshev_eval(int max_moment, int grid_size) : M_(max_moment + max_moment % 2), G_(grid_size), angle_grid(grid_size), lobatto_grid(grid_size),
                                                    chebt_cache(size_t(M_) * grid_size) {
        for (int i = 0; i < G_; i++) {
            angle_grid[i] = (i == G_ - 1) ? 1.0 : double(i) / double(G_ - 1);
            lobatto_grid[i] = -std::cos(M_PI * angle_grid[i]);
            for (int k = 0; k < M_; k++) chebt_cache[size_t(k) * G_ + i] = std::cos(k * std::acos(lobatto_grid[i]));
        }
    }
    template <typename F>
    double moment_f(const F& op, int order) const {
        std::vector<double> vals(G_);
        for (int i = 0; i < G_; ++i) vals[i] = op(lobatto_grid[i]);
        return moment(vals, order);
    }
    double moment(const std::vector<double>& in, int order) const {
        double s = 0.0;
        for (int i = 0; i < G_ - 1; ++i)
            s += (in[i + 1] * chebt_cache[size_t(order) * G_ + i + 1] + in[i] * chebt_cache[size_t(order) * G_ + i]) * (angle_grid[i + 1] - angle_grid[i]);
        return s * 0.5;
    }

protected:
    int M_, G_;
    std::vector<double> angle_grid, lobatto_grid, chebt_cache;
};
}  // namespace chebyshev

struct config_params {
    double beta, U, mu_c, mu_f;
    std::vector<double> W;
    bool operator==(const config_params& rhs) const {
        double tol = std::numeric_limits<double>::epsilon();
        return (std::abs(beta - rhs.beta) < tol && std::abs(U - rhs.U) < tol && std::abs(mu_c - rhs.mu_c) < tol && std::abs(mu_f - rhs.mu_f) < tol);
    }
};

struct ed_cache {
    enum status_eval { empty, spectrum, full };
    typedef std::vector<double> real_array_t;
    status_eval status = empty;
    real_array_t cached_spectrum, cached_exp, cached_fermi;
    std::vector<double> cached_evecs;  // column-major N x N
    double logZ = 0.0;
};

struct chebyshev_cache {
    enum status_eval { empty, logz };
    status_eval status = empty;
    double e_max = 0, e_min = 0, a = 0, b = 0;
    std::vector<double> moments;
    double logZ = 0.0;
    // not in the reference: the trace-sum record of fkmc_logz_kpm_batched_local and the configuration it belongs to.  It survives
    // reset_cache / calc_hamiltonian, so that a proposal (a copy of the current configuration with one or two sites changed) is
    // re-evaluated locally against the configuration it was copied from.
    std::vector<double> kpm_state;
    std::vector<int32_t> kpm_state_f;
};

/// include/fk_mc/configuration.hpp:52-87
struct configuration_t {
    typedef std::vector<double> real_array_t;
    typedef std::vector<int32_t> int_array_t;

    configuration_t(const abstract_lattice& lattice, double beta, double U, double mu_c, double mu_f, std::vector<double> W = {})
        : lattice_(lattice), params_(config_params({beta, U, mu_c, mu_f, W})), f_config_(lattice.volume(), 0) {}
    configuration_t(const configuration_t&) = default;
    configuration_t& operator=(const configuration_t& rhs) {
        f_config_ = rhs.f_config_;
        ed_data_ = rhs.ed_data_;
        cheb_data_ = rhs.cheb_data_;
        if (!(params_ == rhs.params_)) throw(std::logic_error("Mismatched parameters in config assignment"));  // configuration.cpp:38
        return *this;
    }
    size_t get_nf() const { return size_t(std::accumulate(f_config_.begin(), f_config_.end(), 0)); }
    /// src/configuration.cpp:47-56
    void randomize_f(random_generator& rnd, size_t nf = 0) {
        std::uniform_int_distribution<> distr(0, int(lattice_.volume()) - 1);
        if (!nf) nf = distr(rnd);
        std::fill(f_config_.begin(), f_config_.end(), 0);
        for (size_t i = 0; i < nf; ++i) {
            size_t ind = distr(rnd);
            while (f_config_[ind] == 1) ind = distr(rnd);
            f_config_[ind] = 1;
        }
    }
    /// src/configuration.cpp:79-91: the Hamiltonian itself is assembled on the device from f_config_; this only resets the caches
    void calc_hamiltonian() { reset_cache(); }
    void reset_cache() { ed_data_.status = ed_cache::empty; cheb_data_.status = chebyshev_cache::empty; }
    /// src/configuration.cpp:208-246
    void calc_ed(bool calc_evecs = false) {
        if ((ed_data_.status == ed_cache::spectrum && !calc_evecs) || (ed_data_.status == ed_cache::full && calc_evecs)) return;
        const size_t N = lattice_.msize();
        ed_data_.cached_spectrum.resize(N);
        ed_data_.cached_exp.resize(N);
        ed_data_.cached_fermi.resize(N);
        if (calc_evecs) {
            ed_data_.cached_evecs.resize(N * N);
            fkmc_check(fkmc_eigh_batched(lattice_.ctx(), f_config_.data(), 1, params_.U, params_.mu_c, params_.beta, ed_data_.cached_spectrum.data(),
                                         ed_data_.cached_evecs.data(), &ed_data_.logZ), lattice_.ctx());
            for (size_t i = 0; i < N; ++i) {
                ed_data_.cached_exp[i] = std::exp(params_.beta * ed_data_.cached_spectrum[i]);
                ed_data_.cached_fermi[i] = 1.0 / (1.0 + ed_data_.cached_exp[i]);
            }
            ed_data_.status = ed_cache::full;
            return;
        }
        fkmc_check(fkmc_logz_ed_batched(lattice_.ctx(), f_config_.data(), 1, params_.U, params_.mu_c, params_.beta, ed_data_.cached_spectrum.data(),
                                        &ed_data_.logZ, ed_data_.cached_exp.data(), ed_data_.cached_fermi.data()), lattice_.ctx());
        ed_data_.status = ed_cache::spectrum;
    }
    /// src/configuration.cpp:94-205
    void calc_chebyshev(const chebyshev::chebyshev_eval& cheb) {
        if (int(cheb_data_.status) >= int(chebyshev_cache::logz)) return;
        double ab[4];
        cheb_data_.moments.resize(cheb.cheb_size());
        const bool have_ref = cheb_data_.kpm_state.size() == FKMC_KPM_STATE_DOUBLES && cheb_data_.kpm_state_f.size() == f_config_.size();
        std::vector<double> state_out(FKMC_KPM_STATE_DOUBLES);
        fkmc_check(fkmc_logz_kpm_batched_local(lattice_.ctx(), f_config_.data(), have_ref ? cheb_data_.kpm_state_f.data() : nullptr,
                                               have_ref ? cheb_data_.kpm_state.data() : nullptr, 1, params_.U, params_.mu_c, params_.beta, cheb.cheb_size(),
                                               cheb.grid_size(), cheb_data_.moments.data(), ab, &cheb_data_.logZ, state_out.data()), lattice_.ctx());
        cheb_data_.kpm_state.swap(state_out);
        cheb_data_.kpm_state_f.assign(f_config_.begin(), f_config_.end());
        cheb_data_.e_min = ab[0]; cheb_data_.e_max = ab[1]; cheb_data_.a = ab[2]; cheb_data_.b = ab[3];
        cheb_data_.status = chebyshev_cache::logz;
    }
    /// src/configuration.cpp:59-77
    double calc_ff_energy() const {
        if (lattice_.ndim() != 1) return 0;
        double e = 0;
        const int V = int(lattice_.volume());
        for (int i = 0; i < V; ++i) {
            if (!f_config_[i]) continue;
            for (int l = 0; l < int(params_.W.size()); ++l) {
                const int left = (i - l + V) % V, right = (i + l) % V;
                e += params_.W[l] * f_config_[left];
                e += params_.W[l] * f_config_[right] * (l > 0);
            }
        }
        return e;
    }
    const config_params& params() const { return params_; }
    const ed_cache& ed_data() const { return ed_data_; }
    const chebyshev_cache& cheb_data() const { return cheb_data_; }

    const abstract_lattice& lattice_;
    const config_params params_;
    int_array_t f_config_;
    ed_cache ed_data_;
    chebyshev_cache cheb_data_;
};

// ---------------------------------------------------------------- moves (src/moves.cpp) ----
struct move_flip {
    typedef double mc_weight_type;
    double beta;
    configuration_t& config;
    configuration_t new_config;
    random_generator& RND;
    move_flip(double beta, configuration_t& current_config, random_generator& RND_) : beta(beta), config(current_config), new_config(current_config), RND(RND_) {}
    mc_weight_type attempt() {
        std::uniform_int_distribution<> distr(0, int(config.lattice_.volume()) - 1);
        if (config.get_nf() == 0 || config.get_nf() == config.lattice_.msize()) return 0;
        new_config = config;
        size_t from = distr(RND); while (new_config.f_config_[from] == 0) from = distr(RND);
        size_t to = distr(RND); while (new_config.f_config_[to] == 1) to = distr(RND);
        config.calc_ed(false);
        new_config.f_config_[from] = 0;
        new_config.f_config_[to] = 1;
        new_config.calc_hamiltonian();
        new_config.calc_ed(false);
        return std::exp(new_config.ed_data_.logZ - config.ed_data_.logZ);
    }
    mc_weight_type accept() { config = new_config; return 1.0; }
    void reject() {}
};

struct move_randomize : move_flip {
    move_randomize(double beta, configuration_t& c, random_generator& r) : move_flip(beta, c, r) {}
    mc_weight_type attempt() {
        new_config = config;
        new_config.randomize_f(RND);
        new_config.calc_hamiltonian();
        config.calc_ed(false);
        new_config.calc_ed(false);
        const double log_ratio = new_config.ed_data_.logZ - config.ed_data_.logZ;
        const double ff_diff = new_config.calc_ff_energy() - config.calc_ff_energy();
        const double dn = double(new_config.get_nf()) - double(config.get_nf());  // signed
        if (beta * config.params_.mu_f * dn - ff_diff > 2.7182818 - log_ratio) return 1;
        else if (beta * config.params_.mu_f * dn - ff_diff + log_ratio < 0) return 0;
        return std::exp(log_ratio) * std::exp(beta * (config.params_.mu_f * dn - ff_diff));
    }
};

struct move_addremove : move_flip {
    double exp_beta_mu_f;
    move_addremove(double beta, configuration_t& c, random_generator& r) : move_flip(beta, c, r), exp_beta_mu_f(std::exp(beta * config.params_.mu_f)) {}
    mc_weight_type attempt() {
        std::uniform_int_distribution<> distr(0, int(config.lattice_.volume()) - 1);
        new_config = config;
        size_t to = distr(RND);
        new_config.f_config_[to] = 1 - config.f_config_[to];
        config.calc_ed(false);
        new_config.calc_hamiltonian();
        new_config.calc_ed(false);
        const double ff_diff = new_config.calc_ff_energy() - config.calc_ff_energy();
        const double ratio = std::exp(new_config.ed_data_.logZ - config.ed_data_.logZ);
        return (new_config.f_config_[to] ? ratio * exp_beta_mu_f : ratio / exp_beta_mu_f) * std::exp(-beta * ff_diff);
    }
};

// ---------------------------------------------------------------- moves (src/moves_chebyshev.cpp) ----
namespace chebyshev {
struct move_flip {
    typedef double mc_weight_type;
    double beta;
    configuration_t& config;
    configuration_t new_config;
    const chebyshev_eval& cheb_;
    random_generator& RND;
    move_flip(double beta, configuration_t& c, const chebyshev_eval& cheb, random_generator& r) : beta(beta), config(c), new_config(c), cheb_(cheb), RND(r) {}
    mc_weight_type attempt() {
        config.calc_chebyshev(cheb_);
        if (config.get_nf() == 0 || config.get_nf() == config.lattice_.msize()) return 0;
        std::uniform_int_distribution<> distr(0, int(config.lattice_.volume()) - 1);
        new_config = config;
        size_t from = distr(RND); while (new_config.f_config_[from] == 0) from = distr(RND);
        size_t to = distr(RND); while (new_config.f_config_[to] == 1) to = distr(RND);
        new_config.f_config_[from] = 0;
        new_config.f_config_[to] = 1;
        new_config.calc_hamiltonian();
        new_config.calc_chebyshev(cheb_);
        const double ff_diff = new_config.calc_ff_energy() - config.calc_ff_energy();
        return std::exp(new_config.cheb_data_.logZ - config.cheb_data_.logZ - beta * ff_diff);
    }
    mc_weight_type accept() { config = new_config; return 1.0; }
    void reject() {}
};
struct move_randomize : move_flip {
    using move_flip::move_flip;
    mc_weight_type attempt() {
        config.calc_chebyshev(cheb_);
        new_config = config;
        new_config.randomize_f(RND);
        new_config.calc_hamiltonian();
        new_config.calc_chebyshev(cheb_);
        const double log_ratio = new_config.cheb_data_.logZ - config.cheb_data_.logZ;
        const double ff_diff = new_config.calc_ff_energy() - config.calc_ff_energy();
        const double dn = double(new_config.get_nf()) - double(config.get_nf());
        if (beta * config.params_.mu_f * dn - ff_diff > 2.7182818 - log_ratio) return 1;
        else if (beta * config.params_.mu_f * dn - ff_diff + log_ratio < 0) return 0;
        return std::exp(log_ratio) * std::exp(beta * (config.params_.mu_f * dn - ff_diff));
    }
};
struct move_addremove : move_flip {
    double exp_beta_mu_f;
    move_addremove(double beta, configuration_t& c, const chebyshev_eval& cheb, random_generator& r)
        : move_flip(beta, c, cheb, r), exp_beta_mu_f(std::exp(beta * config.params_.mu_f)) {}
    mc_weight_type attempt() {
        std::uniform_int_distribution<> distr(0, int(config.lattice_.volume()) - 1);
        config.calc_chebyshev(cheb_);
        new_config = config;
        size_t to = distr(RND);
        new_config.f_config_[to] = 1 - config.f_config_[to];
        new_config.calc_hamiltonian();
        new_config.calc_chebyshev(cheb_);
        const double ff_diff = new_config.calc_ff_energy() - config.calc_ff_energy();
        const double ratio = std::exp(new_config.cheb_data_.logZ - config.cheb_data_.logZ);
        return (new_config.f_config_[to] ? ratio * exp_beta_mu_f : ratio / exp_beta_mu_f) * std::exp(-beta * ff_diff);
    }
};
}  // namespace chebyshev

/// src/measures/energy.cpp:6-26
struct measure_energy {
    measure_energy(double beta, configuration_t& in, std::vector<double>& energies, std::vector<double>& d2energies, std::vector<double>& c_energies)
        : beta(beta), config(in), _energies(energies), _d2energies(d2energies), _c_energies(c_energies) {}
    void accumulate(double /*sign*/) {
        config.calc_ed(false);
        const auto& spectrum = config.ed_data_.cached_spectrum;
        const auto& exp_e = config.ed_data_.cached_exp;
        _Z++;
        double e_val_c = 0, d2 = 0;
        for (size_t i = 0; i < spectrum.size(); ++i) {
            e_val_c += spectrum[i] / (1.0 + exp_e[i]);
            d2 += spectrum[i] * spectrum[i] / (1.0 + 0.5 * (exp_e[i] + 1. / exp_e[i]));
        }
        const double e_val = e_val_c - double(config.params_.mu_f) * config.get_nf() + config.calc_ff_energy();
        _energies.push_back(e_val);
        _d2energies.push_back(d2 / 2.0);
        _c_energies.push_back(e_val_c);
    }
    double beta;
    configuration_t& config;
    int _Z = 0;
    std::vector<double>&_energies, &_d2energies, &_c_energies;
};

/// src/measures/spectrum.cpp:13-21: running mean of the sorted spectrum
struct measure_spectrum {
    measure_spectrum(configuration_t& in, std::vector<double>& average_spectrum) : config(in), _average_spectrum(average_spectrum) {}
    void accumulate(double /*sign*/) {
        config.calc_ed(false);
        const auto& spectrum = config.ed_data_.cached_spectrum;
        if (_average_spectrum.size() != spectrum.size()) _average_spectrum.assign(spectrum.size(), 0.0);
        for (size_t i = 0; i < spectrum.size(); ++i) _average_spectrum[i] = (_average_spectrum[i] * _Z + spectrum[i]) / (_Z + 1);
        _Z++;
    }
    configuration_t& config;
    int _Z = 0;
    std::vector<double>& _average_spectrum;
};

/// src/measures/spectrum_history.cpp:13-19: history [eigenvalue index][measurement]
struct measure_spectrum_history {
    measure_spectrum_history(configuration_t& in, std::vector<std::vector<double>>& spectrum_history) : config(in), spectrum_history_(spectrum_history) {
        spectrum_history_.resize(config.lattice_.msize());
    }
    void accumulate(double /*sign*/) {
        config.calc_ed(false);
        for (size_t i = 0; i < spectrum_history_.size(); ++i) spectrum_history_[i].push_back(config.ed_data_.cached_spectrum[i]);
    }
    configuration_t& config;
    std::vector<std::vector<double>>& spectrum_history_;
};

/// src/measures/focc_history.cpp:7-12: history [site][measurement]
struct measure_focc {
    measure_focc(configuration_t& in, std::vector<std::vector<double>>& focc_history) : config(in), focc_history_(focc_history) {
        focc_history_.resize(config.lattice_.volume());
    }
    void accumulate(double /*sign*/) {
        for (size_t i = 0; i < focc_history_.size(); ++i) focc_history_[i].push_back(config.f_config_[i]);
    }
    configuration_t& config;
    std::vector<std::vector<double>>& focc_history_;
};

/// include/fk_mc/measures/ipr.hpp:39-56: ||psi_k||_4 / ||psi_k||_2^2 per eigenstate; history [state][measurement].  The eigenvectors stay on
/// the device (fkmc_ipr_batched); the spectrum cache is refreshed from the same solve.
struct measure_ipr {
    measure_ipr(configuration_t& in, std::vector<std::vector<double>>& ipr_vals) : config(in), ipr_vals_(ipr_vals) { ipr_vals_.resize(config.lattice_.msize()); }
    void accumulate(double /*sign*/) {
        const size_t N = config.lattice_.msize();
        std::vector<double> ev(N), ipr(N);
        fkmc_check(fkmc_ipr_batched(config.lattice_.ctx(), config.f_config_.data(), 1, config.params_.U, config.params_.mu_c, config.params_.beta, ev.data(),
                                    ipr.data()), config.lattice_.ctx());
        for (size_t i = 0; i < N; ++i) ipr_vals_[i].push_back(ipr[i]);
    }
    configuration_t& config;
    std::vector<std::vector<double>>& ipr_vals_;
};

/// src/measures/eigenfunctions.cpp:12-18: the full eigenvector matrix (column-major N x N, column k <-> eigenvalue k) per measurement
struct measure_eigenfunctions {
    measure_eigenfunctions(configuration_t& in, std::vector<std::vector<double>>& eigenfunctions) : config(in), eigenfunctions_(eigenfunctions) {}
    void accumulate(double /*sign*/) {
        config.calc_ed(true);
        eigenfunctions_.push_back(config.ed_data_.cached_evecs);
    }
    configuration_t& config;
    std::vector<std::vector<double>>& eigenfunctions_;
};

/// Batched product path: n_chains reference ranks on one GPU (fkmc_chain_*).  chain c == the reference's MPI rank chain0 + c.
struct batched_chains {
    batched_chains(abstract_lattice& lat, int n_chains, const fkmc_chain_params& p) : lat_(lat), n_(n_chains), p_(p) {
        fkmc_check(fkmc_chain_init(lat_.ctx(), n_chains, &p_), lat_.ctx());
    }
    void run_sweeps(int n) { fkmc_check(fkmc_chain_run_sweeps(lat_.ctx(), n), lat_.ctx()); }
    /// observables_t::{energies, d2energies, c_energies} as [measurement][chain]
    int series(std::vector<double>& e, std::vector<double>& d2, std::vector<double>& ec) {
        e.assign(size_t(p_.max_sweeps) * n_, 0.0); d2 = e; ec = e;
        int n_meas = 0;
        fkmc_check(fkmc_chain_get_series(lat_.ctx(), &n_meas, e.data(), d2.data(), ec.data(), nullptr), lat_.ctx());
        e.resize(size_t(n_meas) * n_); d2.resize(e.size()); ec.resize(e.size());
        return n_meas;
    }
    /// per-sweep histories as the C ABI returns them: spectrum_mean [chain][N], the others [measurement][chain][N] (empty when not measured)
    int histories(std::vector<double>& spectrum_mean, std::vector<double>& spectrum_history, std::vector<int32_t>& focc_history, std::vector<double>& ipr_history) {
        const size_t N = lat_.msize(), full = size_t(p_.max_sweeps) * n_ * N;
        const bool exact = !p_.cheb_moves || p_.measure_energy || p_.measure_ipr || p_.measure_eigenfunctions;
        spectrum_mean.assign(exact ? size_t(n_) * N : 0, 0.0);
        spectrum_history.assign(exact && p_.measure_history ? full : 0, 0.0);
        focc_history.assign(p_.measure_history ? full : 0, 0);
        ipr_history.assign(p_.measure_ipr ? full : 0, 0.0);
        int n_meas = 0;
        fkmc_check(fkmc_chain_get_history(lat_.ctx(), &n_meas, spectrum_mean.empty() ? nullptr : spectrum_mean.data(),
                                          spectrum_history.empty() ? nullptr : spectrum_history.data(), focc_history.empty() ? nullptr : focc_history.data(),
                                          ipr_history.empty() ? nullptr : ipr_history.data()), lat_.ctx());
        const size_t used = size_t(n_meas) * n_ * N;
        if (!spectrum_history.empty()) spectrum_history.resize(used);
        if (!focc_history.empty()) focc_history.resize(used);
        if (!ipr_history.empty()) ipr_history.resize(used);
        return n_meas;
    }
    std::vector<int32_t> f_config() {
        std::vector<int32_t> f(size_t(n_) * lat_.volume());
        fkmc_check(fkmc_chain_get_state(lat_.ctx(), f.data(), nullptr, nullptr, nullptr), lat_.ctx());
        return f;
    }
    std::vector<int64_t> naccept() {
        std::vector<int64_t> a(n_);
        fkmc_check(fkmc_chain_get_state(lat_.ctx(), nullptr, nullptr, a.data(), nullptr), lat_.ctx());
        return a;
    }
    abstract_lattice& lat_;
    int n_;
    fkmc_chain_params p_;
};

}  // namespace fk

namespace alps {
typedef double mc_weight_t;

/// Minimal stand-in for alps::params as fk_mc uses it (define<T>(name, default, description), p[name] readable as its type and
/// assignable): CLI / ini parsing is out of scope, the table of names, types and defaults is the reference's.
class params {
public:
    typedef std::variant<long, double, bool, std::string, std::vector<double>> value_t;
    struct ref {
        value_t* v;
        std::string name;
        template <typename T> T as() const {
            if (!v) throw std::logic_error("parameter '" + name + "' is not defined");
            if constexpr (std::is_same<T, std::string>::value) return std::get<std::string>(*v);
            else if constexpr (std::is_same<T, std::vector<double>>::value) return std::get<std::vector<double>>(*v);
            else {
                if (auto q = std::get_if<long>(v)) return T(*q);
                if (auto q = std::get_if<double>(v)) return T(*q);
                if (auto q = std::get_if<bool>(v)) return T(*q);
                throw std::logic_error("parameter '" + name + "' is not numeric");
            }
        }
        template <typename T> operator T() const { return as<T>(); }
        template <typename T> ref& operator=(const T& x) {
            if constexpr (std::is_same<T, bool>::value) *v = x;
            else if constexpr (std::is_integral<T>::value) *v = long(x);
            else if constexpr (std::is_floating_point<T>::value) *v = double(x);
            else if constexpr (std::is_same<T, std::vector<double>>::value) *v = x;
            else *v = std::string(x);
            return *this;
        }
    };
    template <typename T> params& define(const std::string& name, T def, const std::string& descr) {
        if (!vals_.count(name)) (*this)[name] = def;  // an explicitly set value wins over the default, as with alps::params
        descr_[name] = descr;
        return *this;
    }
    ref operator[](const std::string& name) { return ref{&vals_[name], name}; }
    ref operator[](const std::string& name) const {
        auto it = vals_.find(name);
        return ref{it == vals_.end() ? nullptr : const_cast<value_t*>(&it->second), name};
    }
    bool exists(const std::string& name) const { return vals_.count(name) > 0; }
    const std::map<std::string, value_t>& values() const { return vals_; }

private:
    std::map<std::string, value_t> vals_;
    std::map<std::string, std::string> descr_;
};

/// Metropolis engine with the reference's registry and accept test (src/mc_metropolis.cpp:34-61, include/fk_mc/mc_metropolis.hpp)
struct mc_metropolis {
    typedef std::mt19937 random_generator;
    struct move_wrap {
        std::shared_ptr<void> ptr_;
        std::function<mc_weight_t(void)> attempt_, accept_;
        std::function<void(void)> reject_;
    };
    struct measure_wrap {
        std::shared_ptr<void> ptr_;
        std::function<void(mc_weight_t)> accumulate_;
    };
    typedef params parameters_type;
    mc_metropolis(long seed, int rank, long nsweeps, long sweep_len, long ntherm_sweeps)
        : random(seed + rank), rank_(rank), measure_sweeps_(nsweeps), sweep_len_(sweep_len), thermalization_sweeps_(ntherm_sweeps) {}
    /// src/mc_metropolis.cpp:21-32 (define_parameters must have been called): seed SEED + rank
    mc_metropolis(parameters_type const& p, int rank)
        : mc_metropolis(p["SEED"].as<long>(), rank, p["nsweeps"].as<long>(), p["sweep_len"].as<long>(), p["ntherm_sweeps"].as<long>()) {}
    /// src/mc_metropolis.cpp:11-19
    static parameters_type& define_parameters(parameters_type& p) {
        p.define<int>("nsweeps", 1024, "Total number of sweeps (1 sweep = #sweep_len moves + 1 measurement)")
            .define<int>("sweep_len", 16, "Number of moves between subsequent measurements")
            .define<int>("ntherm_sweeps", 1, "How many sweeps to do before start measuring")
            .define<bool>("show_output", true, "Show run progress of mc");
        p["nprocs"] = 1;
        return p;
    }
    long sweep_count() const { return sweep_count_; }
    template <typename Move_t>
    bool add_move(Move_t&& move, std::string name, double move_prob = 1.0) {
        typedef typename std::remove_reference<Move_t>::type m_type;
        auto m = std::make_shared<m_type>(std::forward<Move_t>(move));
        move_wrap w;
        w.ptr_ = m;
        w.attempt_ = [m]() { return m->attempt(); };
        w.accept_ = [m]() { return m->accept(); };
        w.reject_ = [m]() { m->reject(); };
        moves_.push_back(w);
        move_names_.push_back(name);
        move_probs_.push_back(move_prob);
        move_distrib_ = std::discrete_distribution<>(move_probs_.begin(), move_probs_.end());
        return true;
    }
    template <typename Measure_t>
    bool add_measure(Measure_t&& measure, std::string name) {
        typedef typename std::remove_reference<Measure_t>::type m_type;
        auto m = std::make_shared<m_type>(std::forward<Measure_t>(measure));
        measure_wrap w;
        w.ptr_ = m;
        w.accumulate_ = [m](mc_weight_t p) { m->accumulate(p); };
        measures_.emplace(name, w);
        return true;
    }
    void update() {
        if (!moves_.size()) throw std::logic_error("No registered moves");
        for (long m = 0; m < sweep_len_; m++) {
            auto move_index = move_distrib_(random);
            mc_weight_t weight = moves_[move_index].attempt_();
            if (std::abs(weight) > metropolis_distrib_(random)) {
                weight *= moves_[move_index].accept_();
                naccept_++;
                phase_ *= (mc_weight_t(0) < weight) - (weight < mc_weight_t(0));
            } else
                moves_[move_index].reject_();
        }
        sweep_count_++;
    }
    void measure() {
        if (measure_count_ >= thermalization_sweeps_)
            for (auto& m : measures_) m.second.accumulate_(phase_);
        measure_count_++;
    }
    double fraction_completed() const { return double(sweep_count_) / double(measure_sweeps_ + thermalization_sweeps_); }
    /// alps::mcbase::run: do { update(); measure(); } while (fraction_completed() < 1)
    void run() { do { update(); measure(); } while (fraction_completed() < 1.0); }
    double acceptance_rate() const { return double(naccept_) / double(sweep_count_ * sweep_len_); }
    random_generator& rng() { return random; }
    long naccept() const { return naccept_; }

protected:
    std::vector<move_wrap> moves_;
    std::vector<std::string> move_names_;
    std::vector<double> move_probs_;
    std::map<std::string, measure_wrap> measures_;
    random_generator random;
    int rank_;
    long measure_sweeps_, sweep_len_, thermalization_sweeps_;
    long sweep_count_ = 0, measure_count_ = 0, naccept_ = 0;
    std::discrete_distribution<> move_distrib_;
    std::uniform_real_distribution<> metropolis_distrib_ = std::uniform_real_distribution<>(0, 1);
    mc_weight_t phase_ = 1.0;
};
}  // namespace alps

namespace fk {
typedef alps::params parameters_t;

/// include/fk_mc/fk_mc.hpp:13-34: time series of one chain (index-major histories: [index][measurement])
struct observables_t {
    std::vector<double> energies, c_energies, d2energies, nf0, nfpi, spectrum, stiffness;
    std::vector<std::vector<double>> spectrum_history, ipr_history, cond_history, focc_history;
    std::vector<std::vector<double>> eigenfunctions_history;  // [measurement] -> column-major N x N (the reference: std::vector<dense_m>)
    void reserve(int n) { energies.reserve(n); d2energies.reserve(n); c_energies.reserve(n); }
};

/// include/fk_mc/fk_mc.hpp:36-65, fk_mc.hxx:35-125,177-207.  One object is one Markov chain driven step by step through the
/// registered moves and measures (batch of one on the GPU); run_batched() runs `n_chains` such chains -- the reference's MPI ranks
/// rank .. rank + n_chains - 1 -- resident on the GPU (fkmc_chain_*), which is the product path.
template <typename LatticeType>
class fk_mc : public alps::mc_metropolis {
    typedef alps::mc_metropolis base;
    static_assert(!std::is_same<LatticeType, abstract_lattice>::value, "Can't construct mc for an unspecified lattice");

public:
    typedef configuration_t config_t;
    typedef LatticeType lattice_type;
    parameters_t p;
    std::shared_ptr<lattice_type> lattice_ptr;
    std::shared_ptr<configuration_t> config_ptr;
    observables_t observables;

    lattice_type const& lattice() const { return *lattice_ptr; }
    configuration_t const& config() const { return *config_ptr; }
    parameters_t& parameters() { return p; }

    /// fk_mc.hxx:177-207: model and move parameters with the reference's names and defaults (+ those of mc_metropolis); SEED := seed
    static parameters_t& define_parameters(parameters_t& p) {
        base::define_parameters(p);
        p.define<double>("beta", 10.0, "Inverse temperature")
            .define<double>("U", 1.0, "FK U")
            .define<double>("mu_c", 0.5, "Chemical potential of c electrons")
            .define<double>("mu_f", 0.5, "Chemical potential of f electrons");
        p.define<double>("mc_flip", 0.0, "Make flip moves")
            .define<double>("mc_add_remove", 1.0, "Make add/remove moves")
            .define<double>("mc_reshuffle", 0.0, "Make reshuffle moves")
            .define<bool>("cheb_moves", false, "Allow moves using Chebyshev sampling")
            .define<double>("cheb_prefactor", 2.2, "Prefactor for number of Chebyshev polynomials = #ln(Volume)")
            .define<bool>("measure_history", true, "Measure the history")
            .define<int>("Nf_start", 5, "Starting number of f-electrons")
            .define<int>("seed", int(std::random_device()()), "Seed for random number generator")
            .define<bool>("measure_ipr", false, "Measure inverse participation ratio")
            .define<bool>("measure_eigenfunctions", false, "Measure eigenfunctions")
            .define<double>("cond_offset", 0.05, "dos offset from the real axis")
            .define<bool>("measure_stiffness", false, "Measure stiffness/conductivity");
        p["SEED"] = p["seed"].as<long>();
        return p;
    }

    fk_mc(parameters_t const& p_, int rank = 0) : base(p_, rank), p(p_), rank0_(rank) {}

    /// fk_mc.hxx:35-125.  The lattice is taken by reference (it owns a GPU context and is not copyable; the reference copies it).
    void initialize(lattice_type& l, bool randomize_config = true, std::vector<double> wgrid_conductivity = {0.0}) {
        wgrid_cond_ = std::move(wgrid_conductivity);  // used by run_batched when measure_stiffness is set (fk_mc.hxx:101-105 keeps it commented out)
        lattice_ptr = std::shared_ptr<lattice_type>(&l, [](lattice_type*) {});
        const std::vector<double> W = (l.ndim() == 1 && p.exists("W")) ? p["W"].template as<std::vector<double>>() : std::vector<double>();
        config_ptr = std::make_shared<configuration_t>(l, p["beta"], p["U"], p["mu_c"], p["mu_f"], W);
        configuration_t& config = *config_ptr;
        if (randomize_config) config.randomize_f(this->rng(), p["Nf_start"].template as<long>());
        config.calc_hamiltonian();
        const double beta = p["beta"];
        const bool cheb_move = p["cheb_moves"];
        if (cheb_move) {
            int cheb_size = int(std::log(double(l.msize())) * double(p["cheb_prefactor"]));
            cheb_size += cheb_size % 2;
            cheb_ptr_.reset(new chebyshev::chebyshev_eval(cheb_size, std::max(cheb_size * 2, 10)));  // owned for the life of the object
        }
        const double eps = std::numeric_limits<double>::epsilon();
        struct entry { const char* key; const char* name; };
        for (entry m : {entry{"mc_flip", "flip"}, entry{"mc_add_remove", "add_remove"}, entry{"mc_reshuffle", "reshuffle"}}) {
            const double w = p[m.key];
            if (!(w > eps)) continue;
            const std::string nm = m.name;
            if (nm == "flip") {
                if (!cheb_move) this->add_move(move_flip(beta, config, this->rng()), nm, w);
                else this->add_move(chebyshev::move_flip(beta, config, *cheb_ptr_, this->rng()), nm, w);
            } else if (nm == "add_remove") {
                if (!cheb_move) this->add_move(move_addremove(beta, config, this->rng()), nm, w);
                else this->add_move(chebyshev::move_addremove(beta, config, *cheb_ptr_, this->rng()), nm, w);
            } else {
                if (!cheb_move) this->add_move(move_randomize(beta, config, this->rng()), nm, w);
                else this->add_move(chebyshev::move_randomize(beta, config, *cheb_ptr_, this->rng()), nm, w);
            }
        }
        observables.reserve(int(p["nsweeps"].template as<long>()));
        const bool history = p["measure_history"], ipr = p["measure_ipr"];
        if (history && ipr) this->add_measure(measure_ipr(config, observables.ipr_history), "ipr");
        if (bool(p["measure_eigenfunctions"])) this->add_measure(measure_eigenfunctions(config, observables.eigenfunctions_history), "eigenfunctions");
        const bool calc_spectrum = !cheb_move || ipr;
        if (calc_spectrum) {
            this->add_measure(measure_energy(beta, config, observables.energies, observables.d2energies, observables.c_energies), "energy");
            this->add_measure(measure_spectrum(config, observables.spectrum), "spectrum");
            if (history) this->add_measure(measure_spectrum_history(config, observables.spectrum_history), "spectrum_history");
        }
        if (history) this->add_measure(measure_focc(config, observables.focc_history), "focc_history");
    }

    /// The same run for `n_chains` chains at once on the lattice's GPU: chain c is the reference's MPI rank (rank + c) with seed
    /// SEED + rank + c.  Parameters map field by field onto fkmc_chain_params; which measures run follows initialize().
    /// Returns per-chain observables (chain-major, i.e. ready to be concatenated like the reference's gathered ranks).
    std::vector<observables_t> run_batched(lattice_type& l, int n_chains) {
        fkmc_chain_params cp{};
        cp.beta = p["beta"]; cp.U = p["U"]; cp.mu_c = p["mu_c"]; cp.mu_f = p["mu_f"];
        cp.mc_flip = p["mc_flip"]; cp.mc_add_remove = p["mc_add_remove"]; cp.mc_reshuffle = p["mc_reshuffle"];
        cp.cheb_moves = bool(p["cheb_moves"]); cp.cheb_prefactor = p["cheb_prefactor"];
        cp.seed = p["SEED"].template as<long>(); cp.chain0 = rank0_;
        cp.nf_start = int(p["Nf_start"].template as<long>());
        cp.sweep_len = int(p["sweep_len"].template as<long>());
        cp.ntherm_sweeps = int(p["ntherm_sweeps"].template as<long>());
        cp.max_sweeps = int(p["nsweeps"].template as<long>()) + cp.ntherm_sweeps;
        const bool history = p["measure_history"], ipr = p["measure_ipr"];
        cp.measure_ipr = history && ipr;
        cp.measure_energy = !cp.cheb_moves || ipr;
        cp.measure_history = history;
        cp.measure_eigenfunctions = bool(p["measure_eigenfunctions"]);
        if (bool(p["measure_stiffness"]) && l.ndim() >= 2) {
            if (wgrid_cond_.size() > FKMC_MAX_COND_W) throw std::logic_error("at most 32 conductivity frequencies");
            cp.measure_stiffness = 1;
            cp.n_cond_w = int(wgrid_cond_.size());
            cp.cond_offset = p["cond_offset"];
            std::copy(wgrid_cond_.begin(), wgrid_cond_.end(), cp.cond_wgrid);
        }
        // not a reference parameter: rank-one secular re-weighting of the dense moves (fkmc.h: fast_update), same results
        if (p.exists("fast_update")) cp.fast_update = bool(p["fast_update"]) && !cp.cheb_moves && !(cp.mc_reshuffle > 0.0);
        if (l.ndim() == 1 && p.exists("W")) {
            const std::vector<double> W = p["W"].template as<std::vector<double>>();
            if (W.size() > FKMC_MAX_W) throw std::logic_error("at most 8 f-f interaction terms");
            cp.n_W = int(W.size());
            std::copy(W.begin(), W.end(), cp.W);
        }
        batched_chains bc(l, n_chains, cp);
        bc.run_sweeps(cp.max_sweeps);
        std::vector<observables_t> out(n_chains);
        const size_t N = l.msize(), C = n_chains;
        std::vector<double> e, d2, ec, smean, shist, ihist;
        std::vector<int32_t> fhist;
        const int n_meas = cp.measure_energy ? bc.series(e, d2, ec) : 0;
        const int n_hist = bc.histories(smean, shist, fhist, ihist);
        for (size_t c = 0; c < C; ++c) {
            observables_t& o = out[c];
            for (int m = 0; m < n_meas; ++m) {
                o.energies.push_back(e[m * C + c]); o.d2energies.push_back(d2[m * C + c]); o.c_energies.push_back(ec[m * C + c]);
            }
            if (!smean.empty()) o.spectrum.assign(smean.begin() + c * N, smean.begin() + (c + 1) * N);
            auto unpack = [&](auto& src, std::vector<std::vector<double>>& dst) {
                if (src.empty()) return;
                dst.assign(N, std::vector<double>(n_hist));
                for (int m = 0; m < n_hist; ++m)
                    for (size_t i = 0; i < N; ++i) dst[i][m] = double(src[(size_t(m) * C + c) * N + i]);
            };
            unpack(shist, o.spectrum_history);
            unpack(fhist, o.focc_history);
            unpack(ihist, o.ipr_history);
        }
        if (cp.measure_eigenfunctions) {
            std::vector<double> ev(size_t(cp.max_sweeps) * C * N * N);
            int n_ev = 0;
            fkmc_check(fkmc_chain_get_eigenfunctions(l.ctx(), &n_ev, ev.data()), l.ctx());
            for (size_t c = 0; c < C; ++c)
                for (int m = 0; m < n_ev; ++m)
                    out[c].eigenfunctions_history.emplace_back(ev.begin() + (size_t(m) * C + c) * N * N, ev.begin() + (size_t(m) * C + c + 1) * N * N);
        }
        if (cp.measure_stiffness) {
            const size_t nw = size_t(cp.n_cond_w);
            std::vector<double> st(size_t(cp.max_sweeps) * C), cd(size_t(cp.max_sweeps) * C * std::max<size_t>(nw, 1));
            int n_st = 0;
            fkmc_check(fkmc_chain_get_stiffness(l.ctx(), &n_st, st.data(), cd.data()), l.ctx());
            for (size_t c = 0; c < C; ++c) {
                out[c].cond_history.assign(nw, std::vector<double>(n_st));  // frequency-major like stiffness.hpp:86
                for (int m = 0; m < n_st; ++m) {
                    out[c].stiffness.push_back(st[m * C + c]);
                    for (size_t w = 0; w < nw; ++w) out[c].cond_history[w][m] = cd[(size_t(m) * C + c) * nw + w];
                }
            }
        }
        batched_naccept_ = bc.naccept();
        batched_f_ = bc.f_config();
        return out;
    }
    const std::vector<int64_t>& batched_naccept() const { return batched_naccept_; }
    const std::vector<int32_t>& batched_f_config() const { return batched_f_; }

private:
    std::unique_ptr<chebyshev::chebyshev_eval> cheb_ptr_;
    int rank0_ = 0;
    std::vector<double> wgrid_cond_{0.0};
    std::vector<int64_t> batched_naccept_;
    std::vector<int32_t> batched_f_;
};

}  // namespace fk
