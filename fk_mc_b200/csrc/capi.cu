// C-ABI entry points of libfkmc_b200 (declared in include/fkmc.h): context management, the batched
// weight evaluators that stand in for configuration_t::calc_ed / calc_chebyshev, the stage-level
// test entry points and instrumentation.  No CPU fallback: every compute call needs an sm_100 GPU.
#include <algorithm>
#include <cstring>
#include <limits>

#include "common.cuh"

static thread_local std::string g_create_error;

int fkmc_set_error(fkmc_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

// Profiling scopes only RECORD events (no host synchronisation inside the timed region); the
// elapsed times are resolved when the totals are read.
fkmc_prof_scope::fkmc_prof_scope(fkmc_ctx* c, const char* n) : ctx(c), name(n) {
    if (!ctx->profiling) return;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, ctx->stream);
}
fkmc_prof_scope::~fkmc_prof_scope() {
    if (!a) return;
    cudaEventRecord(b, ctx->stream);
    ctx->pending.push_back({name, a, b});
}

static void fkmc_profile_resolve(fkmc_ctx* ctx) {
    for (auto& p : ctx->pending) {
        cudaEventSynchronize(p.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, p.a, p.b);
        auto& e = ctx->prof[p.name];
        e.total_ms += ms;
        e.launches += 1;
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    ctx->pending.clear();
}

namespace {

template <class T>
int dalloc(fkmc_ctx* ctx, T** p, size_t n) {
    if (*p) return FKMC_OK;
    FKMC_CUDA(ctx, cudaMalloc((void**)p, sizeof(T) * (n ? n : 1)));
    return FKMC_OK;
}

}  // namespace

// dense-path workspaces, allocated on first use (a KPM-only context never pays for them)
int fkmc_ensure_dense_ws(fkmc_ctx* ctx) {
    const size_t B = ctx->max_batch, N = ctx->N;
    int rc = 0;
    rc |= dalloc(ctx, &ctx->d_A, B * N * N);
    rc |= dalloc(ctx, &ctx->d_W, B * N * FKMC_SYTRD_NB);
    rc |= dalloc(ctx, &ctx->d_AB, B * N * 9);
    if (ctx->band_bw > 0) rc |= dalloc(ctx, &ctx->d_band, B * fkmc_band_stride((int)N));
    rc |= dalloc(ctx, &ctx->d_d, B * N);
    rc |= dalloc(ctx, &ctx->d_e, B * N);
    rc |= dalloc(ctx, &ctx->d_tau, B * N);
    rc |= dalloc(ctx, &ctx->d_evals, B * N);
    rc |= dalloc(ctx, &ctx->d_aux, 2 * B * N);
    return rc ? FKMC_ERR_CUDA : FKMC_OK;
}

int fkmc_tridiagonalize(fkmc_ctx* ctx, double* d_A, int N, int B, double* d_d, double* d_e) {
    const bool two = ctx->tridiag_mode == 2 && N >= 16 && N <= 1024 && fkmc_sy2sb_smem(N) <= ctx->smem_optin && fkmc_sb2st_smem(N) <= ctx->smem_optin;  // (column-major input: fkmc_launch_sy2sb converts when the tiled kernel applies)
    if (!two) return fkmc_launch_sytrd(ctx, d_A, N, B, d_d, d_e, ctx->d_tau, ctx->d_W);
    int rc = fkmc_launch_sy2sb(ctx, d_A, N, B, ctx->d_AB);
    if (rc) return rc;
    return fkmc_launch_sb2st(ctx, ctx->d_AB, N, B, d_d, d_e);
}

int fkmc_build_tridiag(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double* d_d, double* d_e) {
    const int N = ctx->N;
    if (ctx->tridiag_mode == 2 && fkmc_use_band(ctx) && ctx->d_band && fkmc_sb2st_smem(N) <= ctx->smem_optin) {
        // two-dimensional lattices: folded ordering = band matrix; band -> band (DMMA block bulge chasing) -> tridiagonal
        int rc = fkmc_launch_band_reduce(ctx, d_f, B, U, mu_c, ctx->d_band, ctx->d_AB);
        if (rc) return rc;
        return fkmc_launch_sb2st(ctx, ctx->d_AB, N, B, d_d, d_e);
    }
    if (ctx->tridiag_mode == 2 && fkmc_use_tiled(ctx, N) && fkmc_sy2sb_smem(N) <= ctx->smem_optin && fkmc_sb2st_smem(N) <= ctx->smem_optin) {
        // large matrices: tiled lower-triangular layout, bulk-async dense->band, then band->tridiagonal
        int rc = fkmc_launch_build_h_tiled(ctx, d_f, B, U, mu_c, ctx->d_A);
        if (rc) return rc;
        if ((rc = fkmc_launch_sy2sb_tiled(ctx, ctx->d_A, N, B, ctx->d_AB))) return rc;
        return fkmc_launch_sb2st(ctx, ctx->d_AB, N, B, d_d, d_e);
    }
    int rc = fkmc_launch_build_h(ctx, d_f, B, U, mu_c, ctx->d_A);
    if (rc) return rc;
    return fkmc_tridiagonalize(ctx, ctx->d_A, N, B, d_d, d_e);
}

// Reads (and clears) the device-side non-convergence flag the kernels set when an iteration cap is hit; synchronises the stream.
int fkmc_check_flag(fkmc_ctx* ctx) {
    int flag = 0;
    FKMC_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag) {
        FKMC_CUDA(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
        return fkmc_set_error(ctx, FKMC_ERR_NOCONV, flag & 4 ? "fast update: tracked spectrum failed its consistency check (trace invariant / refresh); rerun with fast_update = 0"
                                                          : (flag & 1 ? "bisection / secular iteration cap hit" : "Lanczos step cap hit before e_min/e_max stagnated"));
    }
    return FKMC_OK;
}

namespace {

int check_flag(fkmc_ctx* ctx) { return fkmc_check_flag(ctx); }

int upload_f(fkmc_ctx* ctx, const int32_t* f, int B) {
    if (!f) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "f is NULL");
    if (B < 1 || B > ctx->max_batch) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "B must be in [1, max_batch]");
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_f, f, sizeof(int32_t) * (size_t)B * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
    return FKMC_OK;
}

}  // namespace

extern "C" {

int fkmc_create(fkmc_ctx** out, int device, int lattice_kind, int L, double t, double tp, int max_batch) {
    if (!out) return FKMC_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fkmc_set_error(nullptr, FKMC_ERR_NO_DEVICE, "no CUDA device: libfkmc_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fkmc_set_error(nullptr, FKMC_ERR_INVALID, "bad device index");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fkmc_set_error(nullptr, FKMC_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return fkmc_set_error(nullptr, FKMC_ERR_NO_DEVICE, "libfkmc_b200 is built for sm_100a only");
    if (max_batch < 1) return fkmc_set_error(nullptr, FKMC_ERR_INVALID, "max_batch must be >= 1");
    fkmc_ctx* ctx = new fkmc_ctx();
    ctx->device = device;
    ctx->kind = lattice_kind;
    ctx->L = L;
    ctx->t = t;
    ctx->tp = tp;
    ctx->max_batch = max_batch;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    int rc = fkmc_build_lattice(ctx);
    if (rc) {
        g_create_error = ctx->err;
        delete ctx;
        return rc;
    }
    auto fail = [&](const char* what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError());
        fkmc_destroy(ctx);
        return FKMC_ERR_CUDA;
    };
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice");
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate");
    ctx->own_stream = true;
    const size_t N = ctx->N, Z = ctx->Z, B = max_batch;
    if (cudaMalloc(&ctx->d_nbr_idx, sizeof(int) * std::max<size_t>(1, Z * N)) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&ctx->d_nbr_val, sizeof(double) * std::max<size_t>(1, Z * N)) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&ctx->d_f, sizeof(int32_t) * B * N) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&ctx->d_out, sizeof(double) * B * 8) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&ctx->d_flag, sizeof(int)) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&ctx->d_moments, sizeof(double) * B * 2 * FKMC_MAX_HALF) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&ctx->d_ab, sizeof(double) * B * 4) != cudaSuccess) return fail("cudaMalloc");
    if (cudaMalloc(&ctx->d_kpm_steps, sizeof(int) * B) != cudaSuccess) return fail("cudaMalloc");
    cudaMemcpy(ctx->d_nbr_idx, ctx->h_nbr_idx.data(), sizeof(int) * Z * N, cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->d_nbr_val, ctx->h_nbr_val.data(), sizeof(double) * Z * N, cudaMemcpyHostToDevice);
    cudaMemset(ctx->d_flag, 0, sizeof(int));
    if (fkmc_band_setup(ctx)) return fail("band setup");
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    if (cudaGetLastError() != cudaSuccess) return fail("context setup");
    *out = ctx;
    return FKMC_OK;
}

int fkmc_destroy(fkmc_ctx* ctx) {
    if (!ctx) return FKMC_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    fkmc_profile_resolve(ctx);
    fkmc_comm_destroy(ctx);
    cudaFree(ctx->d_gather);
    cudaFree(ctx->d_ev_scratch);
    cudaFree(ctx->d_kpm_hop0); cudaFree(ctx->d_ks_io); cudaFree(ctx->d_f_ref); cudaFree(ctx->d_band0); cudaFree(ctx->d_band_perm); cudaFree(ctx->d_band);
    fkmc_chain_free(ctx);
    cudaFree(ctx->d_AB); cudaFree(ctx->d_kpm_steps); cudaFree(ctx->d_s1_scratch);
    cudaFree(ctx->d_nbr_idx); cudaFree(ctx->d_nbr_val); cudaFree(ctx->d_A); cudaFree(ctx->d_W); cudaFree(ctx->d_d);
    cudaFree(ctx->d_e); cudaFree(ctx->d_tau); cudaFree(ctx->d_evals); cudaFree(ctx->d_out); cudaFree(ctx->d_f);
    cudaFree(ctx->d_flag); cudaFree(ctx->d_moments); cudaFree(ctx->d_ab); cudaFree(ctx->d_aux);
    cudaFree(ctx->d_chebt); cudaFree(ctx->d_lobatto); cudaFree(ctx->d_dtheta);
    cudaFree(ctx->d_kpm2_tabi); cudaFree(ctx->d_kpm2_h1); cudaFree(ctx->d_kpm2_off); cudaFree(ctx->d_kpm2_nb);
    cudaFree(ctx->d_tile_mask);
    cudaFree(ctx->d_kpm2_part); cudaFree(ctx->d_kpm2_arrived); cudaFree(ctx->d_kpm2_order);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_f_up) cudaEventDestroy(ctx->ev_f_up);
    if (ctx->ev_ref_up) cudaEventDestroy(ctx->ev_ref_up);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return FKMC_OK;
}

const char* fkmc_last_error(const fkmc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int fkmc_volume(const fkmc_ctx* ctx) { return ctx ? ctx->N : -1; }

int fkmc_set_stream(fkmc_ctx* ctx, void* cuda_stream) {
    if (!ctx) return FKMC_ERR_INVALID;
    if (ctx->own_stream && ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return FKMC_OK;
}

int fkmc_sync(fkmc_ctx* ctx) {
    if (!ctx) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

int fkmc_hopping_dense(const fkmc_ctx* ctx, double* H) {
    if (!ctx || !H) return FKMC_ERR_INVALID;
    const int N = ctx->N;
    std::memset(H, 0, sizeof(double) * (size_t)N * N);
    for (int z = 0; z < ctx->Z; ++z)
        for (int i = 0; i < N; ++i) {
            const int j = ctx->h_nbr_idx[(size_t)z * N + i];
            if (j < N) H[(size_t)i * N + j] += ctx->h_nbr_val[(size_t)z * N + i];
        }
    return FKMC_OK;
}

int fkmc_logz_ed_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta, double* evals, double* logZ,
                         double* cached_exp, double* cached_fermi) {
    if (!ctx) return FKMC_ERR_INVALID;
    int rc = upload_f(ctx, f, B);
    if (rc) return rc;
    if ((rc = fkmc_ensure_dense_ws(ctx))) return rc;
    const size_t N = ctx->N;
    if ((rc = fkmc_build_tridiag(ctx, ctx->d_f, B, U, mu_c, ctx->d_d, ctx->d_e))) return rc;
    double* dexp = cached_exp ? ctx->d_aux : nullptr;
    double* dfer = cached_fermi ? ctx->d_aux + (size_t)ctx->max_batch * N : nullptr;
    if ((rc = fkmc_launch_tridiag_eig(ctx, ctx->d_d, ctx->d_e, ctx->N, B, beta, ctx->d_evals, N, nullptr, 0, ctx->d_out, dexp, dfer)))
        return rc;
    if (evals) FKMC_CUDA(ctx, cudaMemcpyAsync(evals, ctx->d_evals, sizeof(double) * B * N, cudaMemcpyDeviceToHost, ctx->stream));
    if (cached_exp) FKMC_CUDA(ctx, cudaMemcpyAsync(cached_exp, dexp, sizeof(double) * B * N, cudaMemcpyDeviceToHost, ctx->stream));
    if (cached_fermi) FKMC_CUDA(ctx, cudaMemcpyAsync(cached_fermi, dfer, sizeof(double) * B * N, cudaMemcpyDeviceToHost, ctx->stream));
    if (logZ)
        FKMC_CUDA(ctx, cudaMemcpy2DAsync(logZ, sizeof(double), ctx->d_out, 8 * sizeof(double), sizeof(double), B, cudaMemcpyDeviceToHost,
                                         ctx->stream));
    return check_flag(ctx);
}

static int eig_common(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta, double* evals, double* evecs, double* ipr,
                      double* logZ) {
    if (!ctx) return FKMC_ERR_INVALID;
    int rc = upload_f(ctx, f, B);
    if (rc) return rc;
    if ((rc = fkmc_ensure_dense_ws(ctx))) return rc;
    const size_t N = ctx->N;
    if ((rc = fkmc_eigvec_pipeline(ctx, ctx->d_f, B, U, mu_c, beta, ctx->d_evals, ctx->d_out, evecs, ipr, nullptr))) return rc;
    if (evals) FKMC_CUDA(ctx, cudaMemcpyAsync(evals, ctx->d_evals, sizeof(double) * B * N, cudaMemcpyDeviceToHost, ctx->stream));
    if (logZ)
        FKMC_CUDA(ctx, cudaMemcpy2DAsync(logZ, sizeof(double), ctx->d_out, 8 * sizeof(double), sizeof(double), B, cudaMemcpyDeviceToHost,
                                         ctx->stream));
    return check_flag(ctx);
}

int fkmc_eigh_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta, double* evals, double* evecs, double* logZ) {
    if (!evecs) return ctx ? fkmc_set_error(ctx, FKMC_ERR_INVALID, "evecs is NULL") : FKMC_ERR_INVALID;
    return eig_common(ctx, f, B, U, mu_c, beta, evals, evecs, nullptr, logZ);
}

int fkmc_ipr_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta, double* evals, double* ipr) {
    if (!ipr) return ctx ? fkmc_set_error(ctx, FKMC_ERR_INVALID, "ipr is NULL") : FKMC_ERR_INVALID;
    return eig_common(ctx, f, B, U, mu_c, beta, evals, nullptr, ipr, nullptr);
}

int fkmc_logz_kpm_batched(fkmc_ctx* ctx, const int32_t* f, int B, double U, double mu_c, double beta, int M, int G, double* moments,
                          double* ab, double* logZ) {
    if (!ctx) return FKMC_ERR_INVALID;
    if (M < 2 || M % 2 || M > 2 * FKMC_MAX_HALF) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "M must be even and in [2, 32]");
    int rc = upload_f(ctx, f, B);
    if (rc) return rc;
    if ((rc = fkmc_launch_kpm(ctx, ctx->d_f, B, U, mu_c, beta, M, G, ctx->d_moments, ctx->d_ab, ctx->d_out))) return rc;
    if (moments) FKMC_CUDA(ctx, cudaMemcpyAsync(moments, ctx->d_moments, sizeof(double) * (size_t)B * M, cudaMemcpyDeviceToHost, ctx->stream));
    if (ab) FKMC_CUDA(ctx, cudaMemcpyAsync(ab, ctx->d_ab, sizeof(double) * (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (logZ) FKMC_CUDA(ctx, cudaMemcpyAsync(logZ, ctx->d_out, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
    return check_flag(ctx);
}

int fkmc_logz_kpm_batched_local(fkmc_ctx* ctx, const int32_t* f, const int32_t* f_ref, const double* state_ref, int B, double U, double mu_c,
                                double beta, int M, int G, double* moments, double* ab, double* logZ, double* state_out) {
    if (!ctx) return FKMC_ERR_INVALID;
    if (M < 2 || M % 2 || M > 2 * FKMC_MAX_HALF) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "M must be even and in [2, 32]");
    if ((f_ref == nullptr) != (state_ref == nullptr)) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "f_ref and state_ref go together");
    int rc = upload_f(ctx, f, B);
    if (rc) return rc;
    const size_t N = ctx->N, nb = ctx->max_batch;
    if (!ctx->d_ks_io) {
        FKMC_CUDA(ctx, cudaMalloc(&ctx->d_ks_io, sizeof(double) * 2 * nb * FKMC_KPM_STATE));
        FKMC_CUDA(ctx, cudaMalloc(&ctx->d_f_ref, sizeof(int32_t) * nb * N));
        FKMC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        FKMC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_f_up, cudaEventDisableTiming));
        FKMC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ref_up, cudaEventDisableTiming));
    }
    double* ks_in = ctx->d_ks_io;
    double* ks_out = ctx->d_ks_io + nb * FKMC_KPM_STATE;
    if (f_ref) {
        if ((rc = fkmc_kpm_prepare_local(ctx))) return rc;
        // the previous call ended with a stream synchronisation, so nothing still reads d_f_ref / ks_in; the second stream starts after
        // the proposals have arrived (the two uploads would only share the link) and runs under the Lanczos kernel
        FKMC_CUDA(ctx, cudaEventRecord(ctx->ev_f_up, ctx->stream));
        FKMC_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_f_up, 0));
        FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_f_ref, f_ref, sizeof(int32_t) * (size_t)B * N, cudaMemcpyHostToDevice, ctx->copy_stream));
        FKMC_CUDA(ctx, cudaMemcpyAsync(ks_in, state_ref, sizeof(double) * (size_t)B * FKMC_KPM_STATE, cudaMemcpyHostToDevice, ctx->copy_stream));
        FKMC_CUDA(ctx, cudaEventRecord(ctx->ev_ref_up, ctx->copy_stream));
        if (fkmc_kpm2d_applicable(ctx, M)) ctx->kpm_wait_event = ctx->ev_ref_up;  // consumed between the Lanczos and the moments launch
        else FKMC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_ref_up, 0));
        ctx->kpm_f_cur = ctx->d_f_ref;
        ctx->kpm_ks_in = ks_in;
    }
    ctx->kpm_ks_out = state_out ? ks_out : nullptr;
    rc = fkmc_launch_kpm(ctx, ctx->d_f, B, U, mu_c, beta, M, G, ctx->d_moments, ctx->d_ab, ctx->d_out);
    if (ctx->kpm_wait_event) {  // a launch path that did not reach the moments kernel: the stream still has to see the uploads finished
        cudaStreamWaitEvent(ctx->stream, ctx->kpm_wait_event, 0);
        ctx->kpm_wait_event = nullptr;
    }
    const bool wrote_state = ctx->kpm_state_written;
    ctx->kpm_f_cur = nullptr; ctx->kpm_ks_in = nullptr; ctx->kpm_ks_out = nullptr;
    if (rc) return rc;
    if (moments) FKMC_CUDA(ctx, cudaMemcpyAsync(moments, ctx->d_moments, sizeof(double) * (size_t)B * M, cudaMemcpyDeviceToHost, ctx->stream));
    if (ab) FKMC_CUDA(ctx, cudaMemcpyAsync(ab, ctx->d_ab, sizeof(double) * (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (logZ) FKMC_CUDA(ctx, cudaMemcpyAsync(logZ, ctx->d_out, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
    if (state_out) {
        if (wrote_state) {
            FKMC_CUDA(ctx, cudaMemcpyAsync(state_out, ks_out, sizeof(double) * (size_t)B * FKMC_KPM_STATE, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            // lattices served by the full-lattice-vector kernel keep no trace sums: the record says "not valid" and every call is a full one
            for (size_t i = 0; i < (size_t)B * FKMC_KPM_STATE; ++i) state_out[i] = 0.0;
        }
    }
    return check_flag(ctx);
}

int fkmc_energy_from_spectrum(fkmc_ctx* ctx, const double* evals, int B, double beta, double* out3) {
    if (!ctx || !evals || !out3) return FKMC_ERR_INVALID;
    if (B < 1 || B > ctx->max_batch) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "B must be in [1, max_batch]");
    int rc = fkmc_ensure_dense_ws(ctx);
    if (rc) return rc;
    const size_t N = ctx->N;
    FKMC_CUDA(ctx, cudaMemcpyAsync(ctx->d_evals, evals, sizeof(double) * B * N, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = fkmc_launch_energy(ctx, ctx->d_evals, N, nullptr, 0, ctx->N, B, beta, ctx->d_out))) return rc;
    std::vector<double> tmp((size_t)B * 8);
    FKMC_CUDA(ctx, cudaMemcpyAsync(tmp.data(), ctx->d_out, sizeof(double) * B * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int b = 0; b < B; ++b) {
        out3[3 * b + 0] = tmp[8 * b + 1];
        out3[3 * b + 1] = tmp[8 * b + 2];
        out3[3 * b + 2] = tmp[8 * b + 0];
    }
    return FKMC_OK;
}

int fkmc_sytrd_batched(fkmc_ctx* ctx, const double* A, int N, int B, double* d, double* e) {
    if (!ctx || !A || !d || !e || N < 2 || B < 1) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    double *dA = nullptr, *dW = nullptr, *dd = nullptr, *de = nullptr, *dt = nullptr;
    const size_t n = N, b = B;
    FKMC_CUDA(ctx, cudaMalloc(&dA, sizeof(double) * b * n * n));
    FKMC_CUDA(ctx, cudaMalloc(&dW, sizeof(double) * b * n * FKMC_SYTRD_NB));
    FKMC_CUDA(ctx, cudaMalloc(&dd, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMalloc(&de, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMalloc(&dt, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMemcpyAsync(dA, A, sizeof(double) * b * n * n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = fkmc_launch_sytrd(ctx, dA, N, B, dd, de, dt, dW);
    if (!rc) {
        FKMC_CUDA(ctx, cudaMemcpyAsync(d, dd, sizeof(double) * b * n, cudaMemcpyDeviceToHost, ctx->stream));
        FKMC_CUDA(ctx, cudaMemcpy2DAsync(e, sizeof(double) * (n - 1), de, sizeof(double) * n, sizeof(double) * (n - 1), b,
                                         cudaMemcpyDeviceToHost, ctx->stream));
        FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dd); cudaFree(de); cudaFree(dt);
    return rc;
}

int fkmc_sy2sb_batched(fkmc_ctx* ctx, const double* A, int N, int B, double* AB) {
    if (!ctx || !A || !AB || N < 2 || B < 1) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    double *dA = nullptr, *dAB = nullptr;
    const size_t n = N, b = B;
    FKMC_CUDA(ctx, cudaMalloc(&dA, sizeof(double) * b * n * n));
    FKMC_CUDA(ctx, cudaMalloc(&dAB, sizeof(double) * b * n * 9));
    FKMC_CUDA(ctx, cudaMemsetAsync(dAB, 0, sizeof(double) * b * n * 9, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpyAsync(dA, A, sizeof(double) * b * n * n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = fkmc_launch_sy2sb(ctx, dA, N, B, dAB);
    if (!rc) {
        FKMC_CUDA(ctx, cudaMemcpyAsync(AB, dAB, sizeof(double) * b * n * 9, cudaMemcpyDeviceToHost, ctx->stream));
        FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(dA); cudaFree(dAB);
    return rc;
}

int fkmc_sb2st_batched(fkmc_ctx* ctx, const double* AB, int N, int B, double* d, double* e) {
    if (!ctx || !AB || !d || !e || N < 3 || B < 1) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    double *dAB = nullptr, *dd = nullptr, *de = nullptr;
    const size_t n = N, b = B;
    FKMC_CUDA(ctx, cudaMalloc(&dAB, sizeof(double) * b * n * 9));
    FKMC_CUDA(ctx, cudaMalloc(&dd, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMalloc(&de, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMemcpyAsync(dAB, AB, sizeof(double) * b * n * 9, cudaMemcpyHostToDevice, ctx->stream));
    int rc = fkmc_launch_sb2st(ctx, dAB, N, B, dd, de);
    if (!rc) {
        FKMC_CUDA(ctx, cudaMemcpyAsync(d, dd, sizeof(double) * b * n, cudaMemcpyDeviceToHost, ctx->stream));
        FKMC_CUDA(ctx, cudaMemcpy2DAsync(e, sizeof(double) * (n - 1), de, sizeof(double) * n, sizeof(double) * (n - 1), b,
                                         cudaMemcpyDeviceToHost, ctx->stream));
        FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(dAB); cudaFree(dd); cudaFree(de);
    return rc;
}

int fkmc_set_option(fkmc_ctx* ctx, const char* name, int value) {
    if (!ctx || !name) return FKMC_ERR_INVALID;
    // options change which kernels a Metropolis step launches: a captured step graph is stale
    if (ctx->chain.step_graph) {
        cudaStreamSynchronize(ctx->stream);
        cudaGraphExecDestroy(ctx->chain.step_graph);
        ctx->chain.step_graph = nullptr;
    }
    if (std::string(name) == "cuda_graph") {
        ctx->use_graphs = value != 0;
        return FKMC_OK;
    }
    if (std::string(name) == "tridiag") {
        if (value != 1 && value != 2) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "tridiag must be 1 or 2");
        ctx->tridiag_mode = value;
        return FKMC_OK;
    }
    if (std::string(name) == "sb2st_warps") {
        if (value < 0 || value > 16) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sb2st_warps must be in [0, 16]");
        ctx->sb2st_warps = value;
        return FKMC_OK;
    }
    if (std::string(name) == "kpm_generic") {
        ctx->kpm_force_generic = value != 0;
        return FKMC_OK;
    }
    if (std::string(name) == "kpm_generic_schedule") {
        ctx->kpm_no_sched = value != 0;
        return FKMC_OK;
    }
    if (std::string(name) == "kpm_local") {  // 1 (default): Chebyshev moves of the chain engine re-evaluate only the columns near the changed sites
        ctx->kpm_local = value != 0;     // (takes effect at the next fkmc_chain_init)
        return FKMC_OK;
    }
    if (std::string(name) == "kpm_rebase_sweeps") {  // local KPM scheme: sweeps between recomputations from scratch (default 16; 0: never)
        if (value < 0) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "kpm_rebase_sweeps must be >= 0");
        ctx->kpm_rebase = value;
        return FKMC_OK;
    }
    if (std::string(name) == "band_path") {  // 1 (default): eigenvalue-only solves of banded lattice matrices start from the band (sb2sb.cu)
        ctx->band_path = value != 0;
        return FKMC_OK;
    }
    if (std::string(name) == "band_min") {  // smallest N served by the band path (default 256)
        if (value < 16) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "band_min must be >= 16");
        ctx->band_min = value;
        return FKMC_OK;
    }
    if (std::string(name) == "sy2sb_tiled_min") {  // smallest N served by the tiled dense->band kernel (default 256)
        if (value < 64 || value > 1024) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "sy2sb_tiled_min must be in [64, 1024]");
        ctx->tiled_min = value;
        return FKMC_OK;
    }
    if (std::string(name) == "lanczos_max_steps") {  // 0: the kernels' own cap (384)
        if (value < 0) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "lanczos_max_steps must be >= 0");
        ctx->lanczos_cap = value;
        return FKMC_OK;
    }
    if (std::string(name) == "eigvec_v1") {
        ctx->eigvec_v1 = value != 0;
        return FKMC_OK;
    }
    if (std::string(name) == "kpm_v1") {
        ctx->kpm_force_v1 = value != 0;
        return FKMC_OK;
    }
    return fkmc_set_error(ctx, FKMC_ERR_INVALID, std::string("unknown option ") + name);
}

int fkmc_tridiag_eigvals_batched(fkmc_ctx* ctx, const double* d, const double* e, int N, int B, double* evals) {
    if (!ctx || !d || !e || !evals || N < 2 || B < 1) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    double *dd = nullptr, *de = nullptr, *dv = nullptr, *dout = nullptr;
    const size_t n = N, b = B;
    FKMC_CUDA(ctx, cudaMalloc(&dd, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMalloc(&de, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMalloc(&dv, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMalloc(&dout, sizeof(double) * b * 8));
    FKMC_CUDA(ctx, cudaMemsetAsync(de, 0, sizeof(double) * b * n, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpyAsync(dd, d, sizeof(double) * b * n, cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpy2DAsync(de, sizeof(double) * n, e, sizeof(double) * (n - 1), sizeof(double) * (n - 1), b,
                                     cudaMemcpyHostToDevice, ctx->stream));
    int rc = fkmc_launch_tridiag_eig(ctx, dd, de, N, B, 1.0, dv, N, nullptr, 0, dout, nullptr, nullptr);
    if (!rc) {
        FKMC_CUDA(ctx, cudaMemcpyAsync(evals, dv, sizeof(double) * b * n, cudaMemcpyDeviceToHost, ctx->stream));
        rc = check_flag(ctx);
    }
    cudaFree(dd); cudaFree(de); cudaFree(dv); cudaFree(dout);
    return rc;
}

int fkmc_secular_update_batched(fkmc_ctx* ctx, const double* lam, const double* z, const double* rho, int N, int B, double* lam_new) {
    if (!ctx || !lam || !z || !rho || !lam_new || N < 1 || B < 1) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    double *dl = nullptr, *dz = nullptr, *dr = nullptr, *dout = nullptr;
    const size_t n = N, b = B;
    FKMC_CUDA(ctx, cudaMalloc(&dl, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMalloc(&dz, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMalloc(&dr, sizeof(double) * b));
    FKMC_CUDA(ctx, cudaMalloc(&dout, sizeof(double) * b * n));
    FKMC_CUDA(ctx, cudaMemcpyAsync(dl, lam, sizeof(double) * b * n, cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpyAsync(dz, z, sizeof(double) * b * n, cudaMemcpyHostToDevice, ctx->stream));
    FKMC_CUDA(ctx, cudaMemcpyAsync(dr, rho, sizeof(double) * b, cudaMemcpyHostToDevice, ctx->stream));
    int rc = fkmc_launch_secular_only(ctx, N, B, dl, dz, dr, dout);
    if (!rc) {
        FKMC_CUDA(ctx, cudaMemcpyAsync(lam_new, dout, sizeof(double) * b * n, cudaMemcpyDeviceToHost, ctx->stream));
        rc = check_flag(ctx);
    }
    cudaFree(dl); cudaFree(dz); cudaFree(dr); cudaFree(dout);
    return rc;
}

int fkmc_kpm_last_steps(fkmc_ctx* ctx, int B, int32_t* steps) {
    if (!ctx || !steps || B < 1 || B > ctx->max_batch) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaMemcpyAsync(steps, ctx->d_kpm_steps, sizeof(int) * B, cudaMemcpyDeviceToHost, ctx->stream));
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}

int64_t fkmc_launch_count(const fkmc_ctx* ctx) { return ctx ? ctx->launches : -1; }

int fkmc_timer_begin(fkmc_ctx* ctx) {
    if (!ctx) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return FKMC_OK;
}

int fkmc_timer_end(fkmc_ctx* ctx, float* ms) {
    if (!ctx || !ms) return FKMC_ERR_INVALID;
    FKMC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    FKMC_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    FKMC_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return FKMC_OK;
}

int fkmc_profile_enable(fkmc_ctx* ctx, int on) {
    if (!ctx) return FKMC_ERR_INVALID;
    ctx->profiling = on != 0;
    return FKMC_OK;
}

int fkmc_profile_get(fkmc_ctx* ctx, const char* family, double* total_ms, int64_t* launches) {
    if (!ctx || !family) return FKMC_ERR_INVALID;
    fkmc_profile_resolve(ctx);
    auto it = ctx->prof.find(family);
    if (total_ms) *total_ms = it == ctx->prof.end() ? 0.0 : it->second.total_ms;
    if (launches) *launches = it == ctx->prof.end() ? 0 : it->second.launches;
    return FKMC_OK;
}

int fkmc_profile_reset(fkmc_ctx* ctx) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_profile_resolve(ctx);
    ctx->prof.clear();
    return FKMC_OK;
}

}  // extern "C"
