// Device-resident Markov chains: the Metropolis loop of mc_metropolis::update
// (src/mc_metropolis.cpp:34-52), the moves of src/moves.cpp / src/moves_chebyshev.cpp and
// measure_energy (src/measures/energy.cpp:6-26) for many independent chains at once.
//
// One step = propose kernel (move pick + site draws, one warp per chain, lane 0 owns the RNG)
//          -> batched weight evaluation (dense eigensolve or KPM over all chains' proposals)
//          -> accept kernel (weight, u ~ U[0,1), accept test, state commit).
// The RNG consumption order per step is the reference's: [discrete move pick iff > 1 move]
// -> the move's own site draws -> one uniform_real (2 words).
#include <algorithm>
#include <cmath>
#include <limits>
#include <random>

#include "common.cuh"
#include "rng.cuh"

namespace {

struct chain_dev {
    uint32_t* mt;
    int32_t *f_cur, *f_prop;
    double *logz_cur;
    int32_t *cur_slot, *prop_slot;
    int32_t *prop_move, *prop_a, *prop_b;
    int32_t *nf_cur, *nf_prop;
    int64_t* naccept;
    int32_t* acc_flag;  // [n_chains] accept decision of the last step (fast update), may be null
    double *ec_cur, *d2_cur;
    double *eff_cur, *eff_prop;  // calc_ff_energy() of the current / proposed configuration
    double *ks_cur, *ks_prop;    // [n_chains][FKMC_KPM_STATE] trace sums of the local KPM scheme (null: off)
    const double* W;             // f-f interaction W[0..nW) (1-D lattices only; nW = 0 otherwise)
    int nW;
    int V, n_chains;
    int n_moves;
    int move_kind[3];
    double move_cp[3];
    double beta, mu_f, exp_beta_mu_f;
};

__device__ void randomize_f_dev(mt19937_dev& g, int V, int nf, int32_t* f) {
    // src/configuration.cpp:47-56 (f already zeroed)
    if (nf == 0) nf = (int)g.uniform_int((uint32_t)V);
    for (int i = 0; i < nf; ++i) {
        uint32_t ind = g.uniform_int((uint32_t)V);
        while (f[ind] == 1) ind = g.uniform_int((uint32_t)V);
        f[ind] = 1;
    }
}

// configuration_t::calc_ff_energy (src/configuration.cpp:59-77), 1-D only: sum_i f_i sum_l W_l (f_{i-l} + [l > 0] f_{i+l}); whole warp
__device__ double ff_energy_warp(const int32_t* f, int V, const double* W, int nW, int lane) {
    if (nW == 0) return 0.0;
    double e = 0.0;
    for (int i = lane; i < V; i += 32) {
        if (!f[i]) continue;
        for (int l = 0; l < nW; ++l) {
            const int left = (i - l % V + V) % V, right = (i + l) % V;
            e += W[l] * (double)f[left];
            if (l > 0) e += W[l] * (double)f[right];
        }
    }
    return warp_sum(e);
}

__global__ void chain_init_kernel(chain_dev C, int64_t seed, int chain0, int nf_start) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C.n_chains) return;
    int32_t* f = C.f_cur + (size_t)c * C.V;
    for (int i = lane; i < C.V; i += 32) f[i] = 0;
    __syncwarp();
    if (lane == 0) {
        mt19937_dev g(C.mt + (size_t)c * FKMC_MT_WORDS);
        g.seed((uint32_t)(uint64_t)(seed + chain0 + c));
        randomize_f_dev(g, C.V, nf_start, f);
        int nf = 0;
        for (int i = 0; i < C.V; ++i) nf += f[i];
        C.nf_cur[c] = nf;
        C.naccept[c] = 0;
        C.cur_slot[c] = 1;   // the initial evaluation writes slot 0 and then "accepts" it
        C.prop_slot[c] = 0;
        C.prop_move[c] = -2;
    }
    __syncwarp();
    int32_t* fp = C.f_prop + (size_t)c * C.V;
    for (int i = lane; i < C.V; i += 32) fp[i] = f[i];
    const double eff = ff_energy_warp(f, C.V, C.W, C.nW, lane);
    if (lane == 0) { C.nf_prop[c] = C.nf_cur[c]; C.eff_cur[c] = eff; C.eff_prop[c] = eff; }
}

// attempt(): the RNG part.  src/moves.cpp:5-21,35-49,52-67 and twins in src/moves_chebyshev.cpp
__global__ void chain_propose_kernel(chain_dev C) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C.n_chains) return;
    const int V = C.V;
    const int32_t* f = C.f_cur + (size_t)c * V;
    int32_t* fp = C.f_prop + (size_t)c * V;
    int kind = 0, a = -1, b = -1;
    if (lane == 0) {
        mt19937_dev g(C.mt + (size_t)c * FKMC_MT_WORDS);
        int mi = 0;
        if (C.n_moves > 1) {  // std::discrete_distribution draws nothing for a single move
            const double p = g.canonical();
            while (mi < C.n_moves - 1 && C.move_cp[mi] < p) ++mi;  // lower_bound over the cumulative probabilities
        }
        kind = C.move_kind[mi];
        const int nf = C.nf_cur[c];
        if (kind == FKMC_MOVE_ADDREMOVE) {
            a = (int)g.uniform_int((uint32_t)V);
        } else if (kind == FKMC_MOVE_FLIP) {
            if (nf == 0 || nf == V) {
                kind = -1;  // "this move won't work": weight 0, no draws
            } else {
                uint32_t from = g.uniform_int((uint32_t)V);
                while (f[from] == 0) from = g.uniform_int((uint32_t)V);
                uint32_t to = g.uniform_int((uint32_t)V);
                while (f[to] == 1) to = g.uniform_int((uint32_t)V);
                a = (int)from;
                b = (int)to;
            }
        }
    }
    kind = __shfl_sync(0xffffffffu, kind, 0);
    a = __shfl_sync(0xffffffffu, a, 0);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (kind == FKMC_MOVE_RESHUFFLE) {
        for (int i = lane; i < V; i += 32) fp[i] = 0;
        __syncwarp();
        if (lane == 0) {
            mt19937_dev g(C.mt + (size_t)c * FKMC_MT_WORDS);
            randomize_f_dev(g, V, 0, fp);
            int nf = 0;
            for (int i = 0; i < V; ++i) nf += fp[i];
            C.nf_prop[c] = nf;
        }
    } else {
        for (int i = lane; i < V; i += 32) {
            int32_t v = f[i];
            if (kind == FKMC_MOVE_ADDREMOVE && i == a) v = 1 - v;
            if (kind == FKMC_MOVE_FLIP) { if (i == a) v = 0; if (i == b) v = 1; }
            fp[i] = v;
        }
        if (lane == 0) {
            int nf = C.nf_cur[c];
            if (kind == FKMC_MOVE_ADDREMOVE) nf += 1 - 2 * f[a];
            C.nf_prop[c] = nf;
        }
    }
    if (C.nW > 0) {
        __syncwarp();
        const double eff = kind == -1 ? C.eff_cur[c] : ff_energy_warp(fp, V, C.W, C.nW, lane);
        if (lane == 0) C.eff_prop[c] = eff;
    }
    if (lane == 0) { C.prop_move[c] = kind; C.prop_a[c] = a; C.prop_b[c] = b; }
}

struct trace_dev {
    int32_t *move, *a, *b, *acc;
    double *w, *u, *lz;
    long step;  // < 0: no trace
};

// weight formulae + accept test (src/mc_metropolis.cpp:43-50) + accept() (src/moves.cpp:23-28).
// init != 0: commit the evaluated initial configuration without drawing.
__global__ void chain_accept_kernel(chain_dev C, const double* __restrict__ logz_prop, int lz_stride,
                                    const double* __restrict__ ecd2, int ecd2_stride, int init, trace_dev TR) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C.n_chains) return;
    const int V = C.V;
    int accept = 0;
    if (lane == 0) {
        const double lz_new = logz_prop[(size_t)c * lz_stride];
        if (init) {
            accept = 1;
        } else {
            const int kind = C.prop_move[c];
            const double lz_old = C.logz_cur[c];
            const double ff_diff = C.nW > 0 ? C.eff_prop[c] - C.eff_cur[c] : 0.0;  // 0 unless 1-D with W (configuration.cpp:62)
            double w = 0.0;
            if (kind == FKMC_MOVE_ADDREMOVE) {
                const double ratio = exp(lz_new - lz_old);
                const int a = C.prop_a[c];
                const int occupied_after = 1 - C.f_cur[(size_t)c * V + a];
                w = occupied_after ? ratio * C.exp_beta_mu_f : ratio / C.exp_beta_mu_f;
                if (C.nW > 0) w *= exp(-C.beta * ff_diff);  // moves.cpp:63-65
            } else if (kind == FKMC_MOVE_FLIP) {
                w = exp(lz_new - lz_old - C.beta * ff_diff);  // moves_chebyshev.cpp:21-22 (E_ff included uniformly)
            } else if (kind == FKMC_MOVE_RESHUFFLE) {
                const double log_ratio = lz_new - lz_old;
                const double dn = (double)C.nf_prop[c] - (double)C.nf_cur[c];
                if (C.beta * C.mu_f * dn - ff_diff > 2.7182818 - log_ratio) w = 1.0;
                else if (C.beta * C.mu_f * dn - ff_diff + log_ratio < 0) w = 0.0;
                else w = exp(log_ratio) * exp(C.beta * (C.mu_f * dn - ff_diff));
            }
            mt19937_dev g(C.mt + (size_t)c * FKMC_MT_WORDS);
            const double u = g.canonical();
            accept = fabs(w) > u ? 1 : 0;
            if (TR.step >= 0) {
                const size_t o = (size_t)TR.step * C.n_chains + c;
                TR.move[o] = kind; TR.a[o] = C.prop_a[c]; TR.b[o] = C.prop_b[c]; TR.acc[o] = accept;
                TR.w[o] = w; TR.u[o] = u; TR.lz[o] = kind >= 0 ? lz_new : 0.0;
            }
        }
        if (accept) {
            C.logz_cur[c] = lz_new;
            C.nf_cur[c] = C.nf_prop[c];
            C.eff_cur[c] = C.eff_prop[c];
            if (!init) C.naccept[c] += 1;
            if (ecd2) { C.ec_cur[c] = ecd2[(size_t)c * ecd2_stride + 1]; C.d2_cur[c] = ecd2[(size_t)c * ecd2_stride + 2]; }
            const int s = C.prop_slot[c];
            C.cur_slot[c] = s;
            C.prop_slot[c] = 1 - s;
        }
    }
    accept = __shfl_sync(0xffffffffu, accept, 0);
    if (lane == 0 && C.acc_flag) C.acc_flag[c] = init ? 0 : accept;
    if (accept) {
        const int32_t* fp = C.f_prop + (size_t)c * V;
        int32_t* f = C.f_cur + (size_t)c * V;
        for (int i = lane; i < V; i += 32) f[i] = fp[i];
        if (C.ks_cur) {
            for (int i = lane; i < FKMC_KPM_STATE; i += 32) C.ks_cur[(size_t)c * FKMC_KPM_STATE + i] = C.ks_prop[(size_t)c * FKMC_KPM_STATE + i];
        }
    }
}

// measure_energy::accumulate: E = E_c - mu_f N_f + E_ff (E_ff = 0 for D >= 2), d2E, E_c
__global__ void chain_measure_kernel(chain_dev C, const double* __restrict__ ecd2, int stride, double* s_e, double* s_d2,
                                     double* s_ec, int32_t* s_nf, long row) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C.n_chains) return;
    const double ec = ecd2 ? ecd2[(size_t)c * stride + 1] : C.ec_cur[c];
    const double d2 = ecd2 ? ecd2[(size_t)c * stride + 2] : C.d2_cur[c];
    const size_t o = (size_t)row * C.n_chains + c;
    s_ec[o] = ec;
    s_d2[o] = d2;
    s_e[o] = ec - C.mu_f * (double)C.nf_cur[c] + C.eff_cur[c];
    s_nf[o] = C.nf_cur[c];
}

// measure_nf0pi::accumulate (include/fk_mc/measures/fsusc0pi.hpp:36-46): n_f(q=0) = sum f (the nf series) and
// |n_f(q=pi)| = |sum_i (-1)^(sum_d pos_d) f_i| (hypercubic_lattice::FFT_pi, src/lattice/hypercubic.cpp:17-28,78-83).  One warp per chain.
__global__ void chain_nfpi_kernel(chain_dev C, int L, int ndim, int32_t* __restrict__ s_nfpi, long row) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C.n_chains) return;
    const int32_t* f = C.f_cur + (size_t)c * C.V;
    int acc = 0;
    for (int i = lane; i < C.V; i += 32) {
        int idx = i, par = 0;
        for (int d = 0; d < ndim; ++d) { par += idx % L; idx /= L; }
        acc += (par & 1) ? -f[i] : f[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_nfpi[(size_t)row * C.n_chains + c] = acc < 0 ? -acc : acc;
}

__global__ void rng_stream_kernel(uint32_t* state, int64_t seed, int mode, int V, int count, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    mt19937_dev g(state);
    g.seed((uint32_t)(uint64_t)seed);
    for (int i = 0; i < count; ++i)
        out[i] = mode == 0 ? (double)g.next() : (mode == 1 ? (double)g.uniform_int((uint32_t)V) : g.canonical());
}

// measure_spectrum (src/measures/spectrum.cpp:13-21: running mean of the sorted spectrum), measure_spectrum_history
// (src/measures/spectrum_history.cpp:13-19) and measure_focc (src/measures/focc_history.cpp:7-12) for one measured sweep.
// One CTA per chain.  spec points at slot 0 of the spectra; cur_slot == nullptr means "slot 0" (Chebyshev moves: the measurement solve).
__global__ void __launch_bounds__(256) chain_history_kernel(chain_dev C, const double* __restrict__ spec, size_t slot_stride, const int32_t* __restrict__ cur_slot,
                                                          int N, double* __restrict__ spec_mean, long count, double* __restrict__ spec_hist_row,
                                                          int32_t* __restrict__ focc_row) {
    const int c = blockIdx.x;
    const double* sp = spec + (cur_slot ? (size_t)cur_slot[c] * slot_stride : 0) + (size_t)c * N;
    if (spec_mean) {
        double* m = spec_mean + (size_t)c * N;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const double v = sp[i];
            m[i] = (m[i] * (double)count + v) / (double)(count + 1);  // spectrum.cpp:17-19
            if (spec_hist_row) spec_hist_row[(size_t)c * N + i] = v;
        }
    }
    if (focc_row) {
        const int32_t* f = C.f_cur + (size_t)c * C.V;
        for (int i = threadIdx.x; i < C.V; i += blockDim.x) focc_row[(size_t)c * C.V + i] = f[i];
    }
}

template <class T>
int dev_alloc(fkmc_ctx* ctx, T** p, size_t n) {
    FKMC_CUDA(ctx, cudaMalloc((void**)p, sizeof(T) * (n ? n : 1)));
    return FKMC_OK;
}

chain_dev make_dev(fkmc_ctx* ctx) {
    fkmc_chain_state& S = ctx->chain;
    chain_dev C{};
    C.mt = S.mt; C.f_cur = S.f_cur; C.f_prop = S.f_prop; C.logz_cur = S.logz_cur;
    C.cur_slot = S.cur_slot; C.prop_slot = S.prop_slot; C.prop_move = S.prop_move; C.prop_a = S.prop_a; C.prop_b = S.prop_b;
    C.nf_cur = S.nf_cur; C.nf_prop = S.nf_prop; C.naccept = S.naccept; C.ec_cur = S.ec_cur; C.d2_cur = S.d2_cur;
    C.acc_flag = S.fu_acc;
    C.ks_cur = S.ks_cur; C.ks_prop = S.ks_prop;
    C.eff_cur = S.eff_cur; C.eff_prop = S.eff_prop; C.W = S.d_W; C.nW = (ctx->ndim == 1) ? S.p.n_W : 0;
    C.V = ctx->N; C.n_chains = S.n_chains; C.n_moves = S.n_moves;
    for (int i = 0; i < 3; ++i) { C.move_kind[i] = S.move_kind[i]; C.move_cp[i] = S.move_cp[i]; }
    C.beta = S.p.beta; C.mu_f = S.p.mu_f; C.exp_beta_mu_f = std::exp(S.p.beta * S.p.mu_f);  // moves.hpp:49
    return C;
}

// evaluate logZ of f_prop for all chains; returns pointers to (logz, stride) and the E_c/d2E block
int evaluate_proposals(fkmc_ctx* ctx, const double** lz, int* lz_stride, const double** ecd2, int* ecd2_stride, bool full_solve = false) {
    fkmc_chain_state& S = ctx->chain;
    const int C = S.n_chains, N = ctx->N;
    if (S.p.cheb_moves) {
        // local scheme (kpm2d.cu): the trace sums of the current configuration are kept per chain, a one- or two-site proposal only
        // recomputes the columns within M/2 hops of the changed sites; full_solve = the periodic re-base from scratch
        ctx->kpm_f_cur = (S.ks_cur && !full_solve) ? S.f_cur : nullptr;
        ctx->kpm_ks_in = (S.ks_cur && !full_solve) ? S.ks_cur : nullptr;
        ctx->kpm_ks_out = S.ks_prop;
        int rc = fkmc_launch_kpm(ctx, S.f_prop, C, S.p.U, S.p.mu_c, S.p.beta, S.M, S.G, ctx->d_moments, ctx->d_ab, S.logz_prop);
        ctx->kpm_f_cur = nullptr; ctx->kpm_ks_in = nullptr; ctx->kpm_ks_out = nullptr;
        if (rc) return rc;
        *lz = S.logz_prop; *lz_stride = 1; *ecd2 = nullptr; *ecd2_stride = 0;
    } else if (S.fu_vt && !full_solve) {
        // rank-one secular updates of the tracked eigen-decomposition (secular.cu)
        int rc = fkmc_fu_evaluate(ctx);
        if (rc) return rc;
        *lz = ctx->d_out; *lz_stride = 8; *ecd2 = ctx->d_out; *ecd2_stride = 8;
    } else {
        int rc = fkmc_build_tridiag(ctx, S.f_prop, C, S.p.U, S.p.mu_c, ctx->d_d, ctx->d_e);
        if (rc) return rc;
        // spectrum of the proposal goes to the chain's non-current slot: spec[0] + slot * (C*N)
        rc = fkmc_launch_tridiag_eig(ctx, ctx->d_d, ctx->d_e, N, C, S.p.beta, S.spec[0], N, S.prop_slot, (long)C * N, ctx->d_out,
                                     nullptr, nullptr);
        if (rc) return rc;
        *lz = ctx->d_out; *lz_stride = 8; *ecd2 = ctx->d_out; *ecd2_stride = 8;
    }
    return FKMC_OK;
}

}  // namespace

int fkmc_chain_free(fkmc_ctx* ctx) {
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return FKMC_OK;
    cudaFree(S.mt); cudaFree(S.f_cur); cudaFree(S.f_prop); cudaFree(S.logz_cur); cudaFree(S.logz_prop); cudaFree(S.ks_cur); cudaFree(S.ks_prop);
    cudaFree(S.spec[0]); cudaFree(S.cur_slot); cudaFree(S.prop_move); cudaFree(S.prop_a); cudaFree(S.prop_b);
    cudaFree(S.naccept); cudaFree(S.s_energy); cudaFree(S.s_d2energy); cudaFree(S.s_cenergy); cudaFree(S.s_nf); cudaFree(S.s_nfpi);
    cudaFree(S.t_move); cudaFree(S.t_a); cudaFree(S.t_b); cudaFree(S.t_acc); cudaFree(S.t_w); cudaFree(S.t_u); cudaFree(S.t_lz);
    cudaFree(S.nf_cur); cudaFree(S.nf_prop); cudaFree(S.prop_slot); cudaFree(S.ec_cur); cudaFree(S.d2_cur);
    cudaFree(S.eff_cur); cudaFree(S.eff_prop); cudaFree(S.d_W);
    if (S.fu_vt) fkmc_fu_free(ctx);
    if (S.step_graph) cudaGraphExecDestroy(S.step_graph);
    cudaFree(S.spec_mean); cudaFree(S.spec_hist); cudaFree(S.focc_hist); cudaFree(S.ipr_hist); cudaFree(S.ipr_evals); cudaFree(S.eig_hist); cudaFree(S.s_stiff); cudaFree(S.cond_hist); cudaFree(S.d_cond_w);
    S = fkmc_chain_state();
    return FKMC_OK;
}

extern "C" int fkmc_chain_init(fkmc_ctx* ctx, int n_chains, const fkmc_chain_params* p) {
    if (!ctx || !p) return FKMC_ERR_INVALID;
    if (n_chains < 1 || n_chains > ctx->max_batch) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "n_chains must be in [1, max_batch]");
    if (p->sweep_len < 1 || p->max_sweeps < 1) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sweep_len and max_sweeps must be >= 1");
    if (p->nf_start < 0 || p->nf_start > ctx->N) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "nf_start out of range");
    if (p->n_W < 0 || p->n_W > FKMC_MAX_W) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "n_W must be in [0, 8]");
    if (p->measure_stiffness && (p->n_cond_w < 0 || p->n_cond_w > FKMC_MAX_COND_W || (ctx->kind != FKMC_CUBIC2D && ctx->kind != FKMC_CUBIC3D)))
        return fkmc_set_error(ctx, FKMC_ERR_INVALID, "measure_stiffness needs a cubic2d / cubic3d lattice and n_cond_w in [0, 32]");
    if (p->fast_update && (p->cheb_moves || p->mc_reshuffle > std::numeric_limits<double>::epsilon() || ctx->N > 1024 || ctx->N < 2))
        return fkmc_set_error(ctx, FKMC_ERR_INVALID, "fast_update needs exact moves, mc_reshuffle = 0 and 2 <= N <= 1024");
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    fkmc_chain_free(ctx);
    fkmc_chain_state& S = ctx->chain;
    S.p = *p;
    S.n_chains = n_chains;
    // move registry: flip, add_remove, reshuffle in this order, each iff weight > eps (fk_mc.hxx:67-78)
    const double eps = std::numeric_limits<double>::epsilon();
    std::vector<double> probs;
    S.n_moves = 0;
    if (p->mc_flip > eps) { S.move_kind[S.n_moves++] = FKMC_MOVE_FLIP; probs.push_back(p->mc_flip); }
    if (p->mc_add_remove > eps) { S.move_kind[S.n_moves++] = FKMC_MOVE_ADDREMOVE; probs.push_back(p->mc_add_remove); }
    if (p->mc_reshuffle > eps) { S.move_kind[S.n_moves++] = FKMC_MOVE_RESHUFFLE; probs.push_back(p->mc_reshuffle); }
    if (S.n_moves == 0) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "No registered moves");  // mc_metropolis.cpp:35-38
    if (S.n_moves > 1) {
        // cumulative probabilities exactly as std::discrete_distribution builds them
        std::discrete_distribution<> dd(probs.begin(), probs.end());
        std::vector<double> pr = dd.probabilities();
        double cs = 0;
        for (int i = 0; i < S.n_moves; ++i) { cs += pr[i]; S.move_cp[i] = cs; }
        S.move_cp[S.n_moves - 1] = 1.0;
    }
    if (p->cheb_moves) {
        // fk_mc.hxx:60-63
        int cheb_size = int(std::log(double(ctx->N)) * p->cheb_prefactor);
        cheb_size += cheb_size % 2;
        S.M = cheb_size;
        S.G = std::max(cheb_size * 2, 10);
        if (S.M < 2 || S.M > 2 * FKMC_MAX_HALF) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "cheb_prefactor gives an unsupported number of moments");
        int rc = fkmc_prepare_cheb(ctx, S.M, S.G);
        if (rc) return rc;
    }
    // which measures run (fk_mc.hxx:94-124): energy + spectrum whenever an exact spectrum exists per sweep (exact moves, or
    // measure_energy / measure_ipr asked for it); spectrum_history and focc_history with measure_history; ipr with measure_ipr
    const bool exact_measure = !p->cheb_moves || p->measure_energy || p->measure_ipr || p->measure_eigenfunctions;
    if (exact_measure || p->measure_stiffness) {
        int rc = fkmc_ensure_dense_ws(ctx);
        if (rc) return rc;
    }
    const size_t C = n_chains, V = ctx->N;
    const size_t rows = p->max_sweeps, steps = (size_t)p->max_sweeps * p->sweep_len;
    S.active = true;  // from here on fkmc_chain_free releases whatever has been allocated
    int rc = 0;
    rc |= dev_alloc(ctx, &S.mt, C * FKMC_MT_WORDS);
    rc |= dev_alloc(ctx, &S.f_cur, C * V);
    rc |= dev_alloc(ctx, &S.f_prop, C * V);
    rc |= dev_alloc(ctx, &S.logz_cur, C);
    rc |= dev_alloc(ctx, &S.logz_prop, C);
    if (p->cheb_moves && ctx->kpm_local) {
        rc |= dev_alloc(ctx, &S.ks_cur, (size_t)C * FKMC_KPM_STATE);
        rc |= dev_alloc(ctx, &S.ks_prop, (size_t)C * FKMC_KPM_STATE);
        if (!rc) cudaMemsetAsync(S.ks_cur, 0, sizeof(double) * (size_t)C * FKMC_KPM_STATE, ctx->stream);   // valid = 0
        if (!rc) rc |= fkmc_kpm_prepare_local(ctx);
    }
    rc |= dev_alloc(ctx, &S.spec[0], 2 * C * V);
    rc |= dev_alloc(ctx, &S.cur_slot, C);
    rc |= dev_alloc(ctx, &S.prop_move, C);
    rc |= dev_alloc(ctx, &S.prop_a, C);
    rc |= dev_alloc(ctx, &S.prop_b, C);
    rc |= dev_alloc(ctx, &S.naccept, C);
    rc |= dev_alloc(ctx, &S.s_energy, rows * C);
    rc |= dev_alloc(ctx, &S.s_d2energy, rows * C);
    rc |= dev_alloc(ctx, &S.s_cenergy, rows * C);
    rc |= dev_alloc(ctx, &S.s_nf, rows * C);
    rc |= dev_alloc(ctx, &S.s_nfpi, rows * C);
    rc |= dev_alloc(ctx, &S.nf_cur, C);
    rc |= dev_alloc(ctx, &S.nf_prop, C);
    rc |= dev_alloc(ctx, &S.prop_slot, C);
    rc |= dev_alloc(ctx, &S.ec_cur, C);
    rc |= dev_alloc(ctx, &S.d2_cur, C);
    rc |= dev_alloc(ctx, &S.eff_cur, C);
    rc |= dev_alloc(ctx, &S.eff_prop, C);
    rc |= dev_alloc(ctx, &S.d_W, (size_t)FKMC_MAX_W);
    if (exact_measure) {
        rc |= dev_alloc(ctx, &S.spec_mean, C * V);
        if (p->measure_history) rc |= dev_alloc(ctx, &S.spec_hist, rows * C * V);
    }
    if (p->measure_history) rc |= dev_alloc(ctx, &S.focc_hist, rows * C * V);
    if (p->measure_ipr) rc |= dev_alloc(ctx, &S.ipr_hist, rows * C * V);
    if (p->measure_ipr || p->measure_eigenfunctions) rc |= dev_alloc(ctx, &S.ipr_evals, C * V);
    if (p->measure_eigenfunctions) rc |= dev_alloc(ctx, &S.eig_hist, rows * C * V * V);
    if (p->measure_stiffness) {
        rc |= dev_alloc(ctx, &S.s_stiff, rows * C);
        rc |= dev_alloc(ctx, &S.cond_hist, rows * C * (size_t)std::max(p->n_cond_w, 1));
        rc |= dev_alloc(ctx, &S.d_cond_w, (size_t)FKMC_MAX_COND_W);
    }
    if (p->record_trace) {
        rc |= dev_alloc(ctx, &S.t_move, steps * C); rc |= dev_alloc(ctx, &S.t_a, steps * C); rc |= dev_alloc(ctx, &S.t_b, steps * C);
        rc |= dev_alloc(ctx, &S.t_acc, steps * C); rc |= dev_alloc(ctx, &S.t_w, steps * C); rc |= dev_alloc(ctx, &S.t_u, steps * C);
        rc |= dev_alloc(ctx, &S.t_lz, steps * C);
    }
    if (rc) {
        fkmc_chain_free(ctx);
        return FKMC_ERR_CUDA;
    }
    if (p->fast_update && (rc = fkmc_fu_alloc(ctx))) {
        const std::string msg = ctx->err;
        fkmc_chain_free(ctx);
        return fkmc_set_error(ctx, rc, msg);
    }
    S.spec[1] = S.spec[0] + C * V;
    S.sweeps_done = 0;
    S.measured = 0;
    S.spec_count = 0;
    if (S.spec_mean) FKMC_CUDA(ctx, cudaMemsetAsync(S.spec_mean, 0, sizeof(double) * C * V, ctx->stream));
    if (p->measure_stiffness && p->n_cond_w > 0)
        FKMC_CUDA(ctx, cudaMemcpyAsync(S.d_cond_w, p->cond_wgrid, sizeof(double) * p->n_cond_w, cudaMemcpyHostToDevice, ctx->stream));
    if (p->n_W > 0) FKMC_CUDA(ctx, cudaMemcpyAsync(S.d_W, p->W, sizeof(double) * p->n_W, cudaMemcpyHostToDevice, ctx->stream));

    chain_dev D = make_dev(ctx);
    const int blocks = (n_chains * 32 + 127) / 128;
    {
        fkmc_prof_scope ps(ctx, "chain_init");
        chain_init_kernel<<<blocks, 128, 0, ctx->stream>>>(D, p->seed, p->chain0, p->nf_start);
        ctx->launches++;
        FKMC_CUDA(ctx, cudaGetLastError());
    }
    // evaluate the initial configuration (the reference does it lazily inside the first attempt())
    const double *lz, *ecd2;
    int lzs, es;
    rc = evaluate_proposals(ctx, &lz, &lzs, &ecd2, &es, /*full_solve=*/true);
    if (rc) return rc;
    trace_dev TR{};
    TR.step = -1;
    chain_accept_kernel<<<blocks, 128, 0, ctx->stream>>>(D, lz, lzs, ecd2, es, 1, TR);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    if (S.fu_vt && (rc = fkmc_fu_refresh(ctx, /*check=*/0))) return rc;  // eigenvectors of the initial configurations
    return fkmc_check_flag(ctx);  // synchronises; FKMC_ERR_NOCONV when a kernel hit its iteration cap
}

extern "C" int fkmc_chain_run_sweeps(fkmc_ctx* ctx, int n_sweeps) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return fkmc_set_error(ctx, FKMC_ERR_STATE, "fkmc_chain_init has not been called");
    if (S.sweeps_done + n_sweeps > S.p.max_sweeps) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "more sweeps than max_sweeps");
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    chain_dev D = make_dev(ctx);
    const int C = S.n_chains, N = ctx->N;
    const int blocks = (C * 32 + 127) / 128;
    const bool exact_measure = !S.p.cheb_moves || S.p.measure_energy || S.p.measure_ipr;
    // one Metropolis step: propose -> weight evaluation -> accept test (-> eigenvector update of the accepted chains)
    auto step_body = [&](const trace_dev& TR) -> int {
        {
            fkmc_prof_scope ps(ctx, "chain_step");
            chain_propose_kernel<<<blocks, 128, 0, ctx->stream>>>(D);
            ctx->launches++;
        }
        const double *lz, *ecd2;
        int lzs, es;
        int rc = evaluate_proposals(ctx, &lz, &lzs, &ecd2, &es);
        if (rc) return rc;
        {
            fkmc_prof_scope ps(ctx, "chain_step");
            chain_accept_kernel<<<blocks, 128, 0, ctx->stream>>>(D, lz, lzs, ecd2, es, 0, TR);
            ctx->launches++;
        }
        if (S.fu_vt && (rc = fkmc_fu_commit(ctx))) return rc;  // accepted chains: V <- V Q
        return FKMC_OK;
    };
    // The step's launch arguments do not change from step to step (unless a trace row index is recorded), so it is captured once as a
    // CUDA graph and replayed; event profiling needs the individual launches.
    const bool use_graph = ctx->use_graphs && !ctx->profiling && !S.p.record_trace && !S.step_graph_failed && !getenv("FKMC_S1_TIMING");
    for (int sw = 0; sw < n_sweeps; ++sw) {
        for (int m = 0; m < S.p.sweep_len; ++m) {
            trace_dev TR{};
            TR.step = -1;
            if (S.p.record_trace) {
                TR.move = S.t_move; TR.a = S.t_a; TR.b = S.t_b; TR.acc = S.t_acc; TR.w = S.t_w; TR.u = S.t_u; TR.lz = S.t_lz;
                TR.step = S.sweeps_done * S.p.sweep_len + m;
            }
            if (use_graph && !S.step_graph && !S.step_graph_failed) {
                const int64_t l0 = ctx->launches;
                cudaGraph_t g = nullptr;
                bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                if (ok) {
                    const int rc = step_body(TR);
                    const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
                    ok = rc == FKMC_OK && e == cudaSuccess && g != nullptr;
                    if (ok) ok = cudaGraphInstantiate(&S.step_graph, g, 0) == cudaSuccess;
                    if (g) cudaGraphDestroy(g);
                }
                S.step_graph_nodes = (int)(ctx->launches - l0);
                ctx->launches = l0;
                if (!ok) {
                    cudaGetLastError();
                    S.step_graph = nullptr;
                    S.step_graph_failed = true;   // run eagerly from now on
                }
            }
            if (use_graph && S.step_graph) {
                FKMC_CUDA(ctx, cudaGraphLaunch(S.step_graph, ctx->stream));
                ctx->launches += S.step_graph_nodes;
            } else {
                const int rc = step_body(TR);
                if (rc) return rc;
            }
        }
        if (S.fu_vt) {
            const int every = S.p.fu_refresh_sweeps > 0 ? S.p.fu_refresh_sweeps : 64;
            if ((S.sweeps_done + 1) % every == 0) {
                int rc = fkmc_fu_refresh(ctx, /*check=*/1);
                if (rc) return rc;
            }
        }
        // measure(): src/mc_metropolis.cpp:54-61
        if (S.sweeps_done >= S.p.ntherm_sweeps) {
            const double* ecd2 = nullptr;
            int es = 0;
            const double* spec = S.spec[0];        // spectrum of the current configurations ...
            const int32_t* slot = S.cur_slot;      // ... in the chain's current slot (exact moves)
            if (S.p.measure_eigenfunctions) {
                // measure_eigenfunctions (+ measure_ipr): one calc_ed(true) serves both; the spectrum / energy measures hit its cache
                int rc = fkmc_eigvec_pipeline_dev2(ctx, S.f_cur, C, S.p.U, S.p.mu_c, S.p.beta, S.ipr_evals, ctx->d_out,
                                                   S.eig_hist + (size_t)S.measured * C * N * N, nullptr,
                                                   S.p.measure_ipr ? S.ipr_hist + (size_t)S.measured * C * N : nullptr);
                if (rc) return rc;
                if (S.p.cheb_moves) { ecd2 = ctx->d_out; es = 8; spec = S.ipr_evals; slot = nullptr; }
            } else if (S.p.measure_ipr && S.fu_vt) {
                // fast update: the eigenvectors of the current configurations are tracked, the IPR is a reduction over them
                int rc = fkmc_fu_ipr(ctx, S.ipr_hist + (size_t)S.measured * C * N);
                if (rc) return rc;
            } else if (S.p.measure_ipr) {
                // measure_ipr::accumulate (ipr.hpp:39-56): calc_ed(true); the energy / spectrum measures that follow hit its cache
                int rc = fkmc_eigvec_pipeline(ctx, S.f_cur, C, S.p.U, S.p.mu_c, S.p.beta, S.ipr_evals, ctx->d_out, nullptr, nullptr,
                                              S.ipr_hist + (size_t)S.measured * C * N);
                if (rc) return rc;
                if (S.p.cheb_moves) { ecd2 = ctx->d_out; es = 8; spec = S.ipr_evals; slot = nullptr; }
            } else if (S.p.cheb_moves && S.p.measure_energy) {
                // Chebyshev moves never fill ed_data_: measure_energy triggers a fresh exact eigensolve
                int rc = fkmc_build_tridiag(ctx, S.f_cur, C, S.p.U, S.p.mu_c, ctx->d_d, ctx->d_e);
                if (rc) return rc;
                rc = fkmc_launch_tridiag_eig(ctx, ctx->d_d, ctx->d_e, N, C, S.p.beta, S.spec[0], N, nullptr, 0, ctx->d_out, nullptr, nullptr);
                if (rc) return rc;
                ecd2 = ctx->d_out;
                es = 8;
                slot = nullptr;
            }
            if (exact_measure && S.p.measure_energy) {
                fkmc_prof_scope ps(ctx, "chain_step");
                chain_measure_kernel<<<(C + 127) / 128, 128, 0, ctx->stream>>>(D, ecd2, es, S.s_energy, S.s_d2energy, S.s_cenergy, S.s_nf,
                                                                            S.measured);
                ctx->launches++;
            }
            if (S.p.measure_stiffness) {
                // measure_stiffness::accumulate (stiffness.hpp:129-187): its own calc_ed(true) + V^T Jm V on DMMA + Kubo sums
                int rc = fkmc_stiffness_dev(ctx, S.f_cur, C, S.p.U, S.p.mu_c, S.p.beta, S.p.cond_offset, S.p.n_cond_w, S.d_cond_w,
                                            S.s_stiff + (size_t)S.measured * C, S.cond_hist + (size_t)S.measured * C * std::max(S.p.n_cond_w, 1));
                if (rc) return rc;
            }
            {
                fkmc_prof_scope ps(ctx, "chain_step");
                chain_nfpi_kernel<<<blocks, 128, 0, ctx->stream>>>(D, ctx->L, ctx->ndim, S.s_nfpi, S.measured);
                ctx->launches++;
            }
            if (exact_measure || S.p.measure_history) {
                fkmc_prof_scope ps(ctx, "chain_step");
                chain_history_kernel<<<C, 256, 0, ctx->stream>>>(D, spec, (size_t)C * N, slot, N, exact_measure ? S.spec_mean : nullptr, S.spec_count,
                                                                  S.spec_hist ? S.spec_hist + (size_t)S.measured * C * N : nullptr,
                                                                  S.focc_hist ? S.focc_hist + (size_t)S.measured * C * N : nullptr);
                ctx->launches++;
                if (exact_measure) S.spec_count++;
            }
            S.measured++;
        }
        if (S.ks_cur && ctx->kpm_rebase > 0 && (S.sweeps_done + 1) % ctx->kpm_rebase == 0) {
            // local KPM scheme: fresh trace sums of the current configurations under their own scaling (bounds the rounding that the
            // incremental sums pick up and re-centres the scaling the affected columns are evaluated with)
            ctx->kpm_ks_out = S.ks_cur;
            int rc = fkmc_launch_kpm(ctx, S.f_cur, C, S.p.U, S.p.mu_c, S.p.beta, S.M, S.G, ctx->d_moments, ctx->d_ab, ctx->d_out);
            ctx->kpm_ks_out = nullptr;
            if (rc) return rc;
        }
        S.sweeps_done++;
    }
    FKMC_CUDA(ctx, cudaGetLastError());
    // one host synchronisation per call: a Lanczos / bisection iteration cap hit anywhere in these sweeps is reported here
    return fkmc_check_flag(ctx);
}

static int copy_out(fkmc_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!dst || !bytes) return FKMC_OK;
    FKMC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return FKMC_OK;
}

extern "C" int fkmc_chain_get_series(fkmc_ctx* ctx, int* n_measured, double* energies, double* d2energies, double* c_energies,
                                     int32_t* nf) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no chains");
    const size_t n = (size_t)S.measured * S.n_chains;
    if (n_measured) *n_measured = (int)S.measured;
    int rc = 0;
    if (S.p.measure_energy) {
        rc |= copy_out(ctx, energies, S.s_energy, n * 8);
        rc |= copy_out(ctx, d2energies, S.s_d2energy, n * 8);
        rc |= copy_out(ctx, c_energies, S.s_cenergy, n * 8);
        rc |= copy_out(ctx, nf, S.s_nf, n * 4);
    }
    if (rc) return rc;
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

extern "C" int fkmc_chain_get_fsector(fkmc_ctx* ctx, int* n_measured, int32_t* nf0, int32_t* nfpi) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no chains");
    const size_t n = (size_t)S.measured * S.n_chains;
    if (n_measured) *n_measured = (int)S.measured;
    int rc = 0;
    if (S.p.measure_energy || !S.p.cheb_moves) rc |= copy_out(ctx, nf0, S.s_nf, n * 4);
    else if (nf0) return fkmc_set_error(ctx, FKMC_ERR_STATE, "nf0 is recorded with the energy measure");
    rc |= copy_out(ctx, nfpi, S.s_nfpi, n * 4);
    if (rc) return rc;
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

extern "C" int fkmc_chain_get_state(fkmc_ctx* ctx, int32_t* f, double* logZ, int64_t* naccept, double* spectrum) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no chains");
    const size_t C = S.n_chains, V = ctx->N;
    int rc = 0;
    rc |= copy_out(ctx, f, S.f_cur, C * V * 4);
    rc |= copy_out(ctx, logZ, S.logz_cur, C * 8);
    rc |= copy_out(ctx, naccept, S.naccept, C * 8);
    if (rc) return rc;
    if (spectrum) {
        if (S.p.cheb_moves) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no cached spectrum with Chebyshev moves");
        std::vector<int32_t> slot(C);
        FKMC_CUDA(ctx, cudaMemcpyAsync(slot.data(), S.cur_slot, C * 4, cudaMemcpyDeviceToHost, ctx->stream));
        FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t c = 0; c < C; ++c)
            FKMC_CUDA(ctx, cudaMemcpyAsync(spectrum + c * V, S.spec[slot[c]] + c * V, V * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

extern "C" int fkmc_chain_get_trace(fkmc_ctx* ctx, int* n_steps, int32_t* move, int32_t* site_a, int32_t* site_b, int32_t* accepted,
                                    double* weight, double* u, double* logz_new) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active || !S.p.record_trace) return fkmc_set_error(ctx, FKMC_ERR_STATE, "trace recording is off");
    const size_t n = (size_t)S.sweeps_done * S.p.sweep_len * S.n_chains;
    if (n_steps) *n_steps = (int)(S.sweeps_done * S.p.sweep_len);
    int rc = 0;
    rc |= copy_out(ctx, move, S.t_move, n * 4);
    rc |= copy_out(ctx, site_a, S.t_a, n * 4);
    rc |= copy_out(ctx, site_b, S.t_b, n * 4);
    rc |= copy_out(ctx, accepted, S.t_acc, n * 4);
    rc |= copy_out(ctx, weight, S.t_w, n * 8);
    rc |= copy_out(ctx, u, S.t_u, n * 8);
    rc |= copy_out(ctx, logz_new, S.t_lz, n * 8);
    if (rc) return rc;
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

extern "C" int fkmc_chain_ipr(fkmc_ctx* ctx, double* evals, double* ipr) {
    if (!ctx || !ipr) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no chains");
    int rc = fkmc_ensure_dense_ws(ctx);
    if (rc) return rc;
    const size_t C = S.n_chains, N = ctx->N;
    if ((rc = fkmc_eigvec_pipeline(ctx, S.f_cur, (int)C, S.p.U, S.p.mu_c, S.p.beta, ctx->d_evals, ctx->d_out, nullptr, ipr, nullptr))) return rc;
    if (evals) FKMC_CUDA(ctx, cudaMemcpyAsync(evals, ctx->d_evals, sizeof(double) * C * N, cudaMemcpyDeviceToHost, ctx->stream));
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

extern "C" int fkmc_chain_get_history(fkmc_ctx* ctx, int* n_measured, double* spectrum_mean, double* spectrum_history, int32_t* focc_history,
                                      double* ipr_history) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no chains");
    const size_t C = S.n_chains, N = ctx->N, n = (size_t)S.measured * C * N;
    if (n_measured) *n_measured = (int)S.measured;
    if ((spectrum_mean && !S.spec_mean) || (spectrum_history && !S.spec_hist) || (focc_history && !S.focc_hist) || (ipr_history && !S.ipr_hist))
        return fkmc_set_error(ctx, FKMC_ERR_STATE, "this history is not being measured (measure_history / measure_ipr / exact spectrum)");
    int rc = 0;
    rc |= copy_out(ctx, spectrum_mean, S.spec_mean, C * N * 8);
    rc |= copy_out(ctx, spectrum_history, S.spec_hist, n * 8);
    rc |= copy_out(ctx, focc_history, S.focc_hist, n * 4);
    rc |= copy_out(ctx, ipr_history, S.ipr_hist, n * 8);
    if (rc) return rc;
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

extern "C" int fkmc_chain_get_eigenfunctions(fkmc_ctx* ctx, int* n_measured, double* evecs) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no chains");
    if (!S.eig_hist) return fkmc_set_error(ctx, FKMC_ERR_STATE, "measure_eigenfunctions is off");
    if (n_measured) *n_measured = (int)S.measured;
    const size_t N = ctx->N;
    int rc = copy_out(ctx, evecs, S.eig_hist, (size_t)S.measured * S.n_chains * N * N * 8);
    if (rc) return rc;
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

extern "C" int fkmc_chain_get_stiffness(fkmc_ctx* ctx, int* n_measured, double* stiffness, double* cond) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no chains");
    if (!S.s_stiff) return fkmc_set_error(ctx, FKMC_ERR_STATE, "measure_stiffness is off");
    if (n_measured) *n_measured = (int)S.measured;
    const size_t n = (size_t)S.measured * S.n_chains;
    int rc = copy_out(ctx, stiffness, S.s_stiff, n * 8);
    if (S.p.n_cond_w > 0) rc |= copy_out(ctx, cond, S.cond_hist, n * S.p.n_cond_w * 8);
    if (rc) return rc;
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

extern "C" int fkmc_chain_series_dev(fkmc_ctx* ctx, void** energies, void** d2energies, void** c_energies, int* ld) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no chains");
    if (energies) *energies = S.s_energy;
    if (d2energies) *d2energies = S.s_d2energy;
    if (c_energies) *c_energies = S.s_cenergy;
    if (ld) *ld = S.n_chains;
    return FKMC_OK;
}

extern "C" int fkmc_rng_stream(fkmc_ctx* ctx, int64_t seed, int mode, int V, int count, double* out) {
    if (!ctx || !out || count < 0) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t* st = nullptr;
    double* d = nullptr;
    FKMC_CUDA(ctx, cudaMalloc(&st, sizeof(uint32_t) * FKMC_MT_WORDS));
    FKMC_CUDA(ctx, cudaMalloc(&d, sizeof(double) * (count ? count : 1)));
    rng_stream_kernel<<<1, 32, 0, ctx->stream>>>(st, seed, mode, V, count, d);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    FKMC_CUDA(ctx, cudaMemcpyAsync(out, d, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(st);
    cudaFree(d);
    return FKMC_OK;
}
