// ORACLE (test infrastructure only) -- flat C entry points for ctypes (tests/, bench.py
// cpu_baseline / --impl reference, __graft_entry__.smoke()).  Never linked into the product.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <thread>

#include "oracle.hpp"

using namespace orc;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH                      \
    }                                  \
    catch (std::exception & e) {       \
        g_err = e.what();              \
        return -1;                     \
    }                                  \
    return 0;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_lattice_size(int kind, int L) {
    int d = (kind == CUBIC1D) ? 1 : (kind == CUBIC3D ? 3 : 2);
    int n = 1;
    for (int i = 0; i < d; ++i) n *= L;
    return n;
}

// dense literal hopping matrix, row-major: H[i*N + j] = hopping_m(i, j)
int orc_hopping_dense(int kind, int L, double t, double tp, double* H) {
    ORC_TRY
    lattice l = make_lattice(kind, L, t, tp);
    std::memset(H, 0, sizeof(double) * (size_t)l.N * l.N);
    for (int i = 0; i < l.N; ++i)
        for (auto& e : l.rows[i]) H[(size_t)i * l.N + e.first] += e.second;
    ORC_CATCH
}

int orc_index_to_pos(int kind, int L, int index, int* pos) {
    ORC_TRY
    lattice l = make_lattice(kind, L, 1.0, 1.0);
    auto p = l.index_to_pos(index);
    for (int i = 0; i < l.ndim; ++i) pos[i] = p[i];
    ORC_CATCH
}

int orc_randomize_f(long seed, int V, int nf, int* f, int* words_after) {
    ORC_TRY
    random_generator r(seed);
    std::vector<int> ff;
    randomize_f(r, V, (size_t)nf, ff);
    std::memcpy(f, ff.data(), sizeof(int) * V);
    if (words_after) *words_after = (int)r();  // next raw word, lets tests pin the stream position
    ORC_CATCH
}

// raw libstdc++ streams for pinning the device RNG: mode 0 = raw words, 1 = uniform_int(0,V-1), 2 = uniform_real(0,1)
int orc_rng_stream(long seed, int mode, int V, int count, double* out) {
    ORC_TRY
    random_generator r(seed);
    std::uniform_int_distribution<> di(0, V > 0 ? V - 1 : 0);
    std::uniform_real_distribution<> dr(0, 1);
    for (int i = 0; i < count; ++i) out[i] = mode == 0 ? double(r()) : (mode == 1 ? double(di(r)) : dr(r));
    ORC_CATCH
}

double orc_ff_energy(int V, const int* f, int nW, const double* W) {
    return calc_ff_energy(1, std::vector<int>(f, f + V), std::vector<double>(W, W + nW));
}

// A col-major n*n (lower triangle read)
int orc_eigh(int n, const double* A, double* evals, double* evecs) { return eigh_lower(n, A, evals, evecs); }

int orc_tridiag(int n, const double* A, double* diag, double* sub) {
    ORC_TRY
    std::vector<double> a(A, A + (size_t)n * n), d, s;
    tridiagonalize_lower(n, a, d, s, nullptr);
    std::memcpy(diag, d.data(), sizeof(double) * n);
    std::memcpy(sub, s.data(), sizeof(double) * (n - 1));
    ORC_CATCH
}

int orc_tridiag_eig(int n, const double* diag, const double* sub, double* evals) {
    ORC_TRY
    std::vector<double> d(diag, diag + n), s(sub, sub + n - 1);
    tridiag_ql_implicit(n, d.data(), s.data(), nullptr);
    std::sort(d.begin(), d.end());
    std::memcpy(evals, d.data(), sizeof(double) * n);
    ORC_CATCH
}

int orc_calc_ed(int kind, int L, double t, double tp, const int* f, double U, double mu_c, double beta, int want_evecs,
                double* spectrum, double* cexp, double* cfermi, double* evecs, double* logZ) {
    ORC_TRY
    lattice l = make_lattice(kind, L, t, tp);
    ed_result r;
    calc_ed(l, std::vector<int>(f, f + l.N), U, mu_c, beta, want_evecs != 0, r);
    if (spectrum) std::memcpy(spectrum, r.spectrum.data(), sizeof(double) * l.N);
    if (cexp) std::memcpy(cexp, r.cached_exp.data(), sizeof(double) * l.N);
    if (cfermi) std::memcpy(cfermi, r.cached_fermi.data(), sizeof(double) * l.N);
    if (evecs && want_evecs) std::memcpy(evecs, r.evecs.data(), sizeof(double) * (size_t)l.N * l.N);
    if (logZ) *logZ = r.logZ;
    ORC_CATCH
}

double orc_logz_from_spectrum(int n, const double* spectrum, double beta) {
    return logz_from_spectrum(std::vector<double>(spectrum, spectrum + n), beta, nullptr, nullptr);
}

void orc_cheb_sizes(int msize, double prefactor, int* M, int* G) { cheb_sizes(msize, prefactor, *M, *G); }

// chebt: [M][G], lobatto: [G]
int orc_cheb_table(int max_moment, int G, double* chebt, double* lobatto, double* angle) {
    ORC_TRY
    chebyshev_eval c(max_moment, G);
    if (chebt) std::memcpy(chebt, c.chebt.data(), sizeof(double) * c.chebt.size());
    if (lobatto) std::memcpy(lobatto, c.lobatto_grid.data(), sizeof(double) * G);
    if (angle) std::memcpy(angle, c.angle_grid.data(), sizeof(double) * G);
    ORC_CATCH
}

// moment of tabulated values vals[G] (vals[i] = F(lobatto[i]))
double orc_cheb_moment(int max_moment, int G, const double* vals, int order) {
    chebyshev_eval c(max_moment, G);
    return c.moment(std::vector<double>(vals, vals + G), order);
}

// out5 = {e_min, e_max, a, b, logZ}
int orc_calc_chebyshev(int kind, int L, double t, double tp, const int* f, double U, double mu_c, double beta, int M, int G,
                       int emode, int prune, double* moments, double* out5, int* lanczos_steps) {
    ORC_TRY
    lattice l = make_lattice(kind, L, t, tp);
    chebyshev_eval c(M, G);
    cheb_result r;
    calc_chebyshev(l, std::vector<int>(f, f + l.N), U, mu_c, beta, c, emode, prune != 0, r);
    if (moments) std::memcpy(moments, r.moments.data(), sizeof(double) * r.moments.size());
    if (out5) { out5[0] = r.e_min; out5[1] = r.e_max; out5[2] = r.a; out5[3] = r.b; out5[4] = r.logZ; }
    if (lanczos_steps) *lanczos_steps = r.lanczos_steps;
    ORC_CATCH
}

int orc_measure_ipr(int n, const double* evecs, double* ipr) {
    ORC_TRY
    std::vector<double> out;
    measure_ipr(n, std::vector<double>(evecs, evecs + (size_t)n * n), out);
    std::memcpy(ipr, out.data(), sizeof(double) * n);
    ORC_CATCH
}

// Full chain.  Trace arrays (nullable) have (nsweeps+ntherm)*sweep_len entries; series have nsweeps.
int orc_mc_run(const mc_params* p, int rank, int* tr_move, int* tr_site_a, int* tr_site_b, int* tr_acc, double* tr_weight,
               double* tr_u, double* tr_logz_new, double* energies, double* d2energies, double* c_energies, double* spectrum_avg,
               int* f_final, long* naccept, double* logz_final, double* ipr_history) {
    ORC_TRY
    mc_result res;
    mc_trace tr;
    const bool want_trace = tr_move || tr_site_a || tr_weight || tr_u || tr_acc || tr_logz_new;
    mc_run(*p, rank, res, want_trace ? &tr : nullptr);
    const size_t ns = tr.move.size();
    if (tr_move) std::memcpy(tr_move, tr.move.data(), sizeof(int) * ns);
    if (tr_site_a) std::memcpy(tr_site_a, tr.site_a.data(), sizeof(int) * ns);
    if (tr_site_b) std::memcpy(tr_site_b, tr.site_b.data(), sizeof(int) * ns);
    if (tr_acc) std::memcpy(tr_acc, tr.accepted.data(), sizeof(int) * ns);
    if (tr_weight) std::memcpy(tr_weight, tr.weight.data(), sizeof(double) * ns);
    if (tr_u) std::memcpy(tr_u, tr.u.data(), sizeof(double) * ns);
    if (tr_logz_new) std::memcpy(tr_logz_new, tr.logz_new.data(), sizeof(double) * ns);
    if (energies && !res.energies.empty()) std::memcpy(energies, res.energies.data(), sizeof(double) * res.energies.size());
    if (d2energies && !res.d2energies.empty()) std::memcpy(d2energies, res.d2energies.data(), sizeof(double) * res.d2energies.size());
    if (c_energies && !res.c_energies.empty()) std::memcpy(c_energies, res.c_energies.data(), sizeof(double) * res.c_energies.size());
    if (spectrum_avg) std::memcpy(spectrum_avg, res.spectrum_avg.data(), sizeof(double) * res.spectrum_avg.size());
    if (f_final) std::memcpy(f_final, res.f_final.data(), sizeof(int) * res.f_final.size());
    if (naccept) *naccept = res.naccept;
    if (logz_final) *logz_final = res.logz_final;
    if (ipr_history)
        for (size_t m = 0; m < res.ipr_history.size(); ++m)
            std::memcpy(ipr_history + m * res.ipr_history[m].size(), res.ipr_history[m].data(), sizeof(double) * res.ipr_history[m].size());
    ORC_CATCH
}

// per-sweep histories of one chain: spectrum_history [nsweeps][N], focc_history [nsweeps][V] (either may be null)
int orc_mc_histories(const mc_params* p, int rank, double* spectrum_history, int* focc_history) {
    ORC_TRY
    mc_result res;
    mc_run(*p, rank, res, nullptr);
    if (spectrum_history)
        for (size_t m = 0; m < res.spectrum_history.size(); ++m)
            std::memcpy(spectrum_history + m * res.spectrum_history[m].size(), res.spectrum_history[m].data(), sizeof(double) * res.spectrum_history[m].size());
    if (focc_history)
        for (size_t m = 0; m < res.focc_history.size(); ++m)
            std::memcpy(focc_history + m * res.focc_history[m].size(), res.focc_history[m].data(), sizeof(int) * res.focc_history[m].size());
    ORC_CATCH
}

// CPU baseline: nthreads independent chains (rank r seeded SEED+r, src/mc_metropolis.cpp:25), exactly
// the reference's MPI layout (ranks never communicate while sampling).  Timed with steady_clock
// around the sampling loop like prog/fk_mc_exec.cpp:149-152.  Returns wall seconds in *seconds.
int orc_bench_chains(const mc_params* p, int nthreads, int rank0, double* seconds, long* naccept_total) {
    ORC_TRY
    std::vector<std::thread> th;
    std::vector<long> acc(nthreads, 0);
    std::vector<std::string> errs(nthreads);
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < nthreads; ++r)
        th.emplace_back([&, r]() {
            try {
                mc_result res;
                mc_run(*p, rank0 + r, res, nullptr);
                acc[r] = res.naccept;
            } catch (std::exception& e) { errs[r] = e.what(); }
        });
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    for (auto& e : errs)
        if (!e.empty()) throw std::logic_error(e);
    *seconds = std::chrono::duration<double>(t1 - t0).count();
    long tot = 0;
    for (long a : acc) tot += a;
    if (naccept_total) *naccept_total = tot;
    ORC_CATCH
}

}  // extern "C"
