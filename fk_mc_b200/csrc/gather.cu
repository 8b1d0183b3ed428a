// End-of-run collective of the Markov chains: the per-chain observable series of every GPU are all-gathered over NCCL
// (NVLink / NVSwitch) straight from the device buffers of the chain engine -- no host round trip, no torch.
//
// Replaces the root-0 gathers of measure_energy::collect_results (src/measures/energy.cpp:32-47: three reduce + three gather of
// `n_meas` doubles per rank).  Here a "rank" of the reference is one chain; GPU g owns chains [g*C, (g+1)*C), so the gathered
// series [measurement][global chain] does not depend on the number of GPUs.
//
// libnccl is resolved at run time (dlopen): the library has no link-time NCCL dependency, single-GPU users never load it, and a
// process that already carries an NCCL (e.g. through torch) shares that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "common.cuh"

namespace {

struct nccl_api {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

nccl_api* load_nccl() {
    static nccl_api api;  // initialised once (C++11 guarantees a thread-safe static); read-only afterwards
    static const bool ok = [] {
        const char* names[] = {getenv("FKMC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return false;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        return api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
    }();
    return ok ? &api : nullptr;
}

const char* nccl_load_error() { return "libnccl could not be loaded or lacks a required symbol (FKMC_NCCL_LIB overrides the library name)"; }

#define FKMC_NCCL(ctx, api, call)                                                                         \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess) return fkmc_set_error(ctx, FKMC_ERR_CUDA, std::string(#call) + ": " + (api)->GetErrorString(r__)); \
    } while (0)

// gathered [rank][m][c] -> [m][rank*C + c] for the three observables at once (grid.y = observable)
__global__ void __launch_bounds__(256) reorder_series_kernel(const double* __restrict__ in, double* __restrict__ out, int nranks, int n_meas, int C) {
    const size_t per = (size_t)nranks * n_meas * C;
    const double* src = in + (size_t)blockIdx.y * per;
    double* dst = out + (size_t)blockIdx.y * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C), m = (int)((i / C) % n_meas), r = (int)(i / ((size_t)C * n_meas));
        dst[((size_t)m * nranks + r) * C + c] = src[i];
    }
}

}  // namespace

extern "C" int fkmc_nccl_unique_id(void* id128) {
    if (!id128) return FKMC_ERR_INVALID;
    nccl_api* api = load_nccl();
    if (!api) return fkmc_set_error(nullptr, FKMC_ERR_STATE, nccl_load_error());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return fkmc_set_error(nullptr, FKMC_ERR_CUDA, "ncclGetUniqueId failed");
    std::memcpy(id128, &id, sizeof(id));
    return FKMC_OK;
}

extern "C" int fkmc_comm_init(fkmc_ctx* ctx, const void* id128, int nranks, int rank) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return FKMC_ERR_INVALID;
    nccl_api* api = load_nccl();
    if (!api) return fkmc_set_error(ctx, FKMC_ERR_STATE, nccl_load_error());
    if (ctx->nccl_comm) return fkmc_set_error(ctx, FKMC_ERR_STATE, "communicator already initialised");
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    FKMC_NCCL(ctx, api, api->CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->nccl_nranks = nranks;
    ctx->nccl_rank = rank;
    return FKMC_OK;
}

extern "C" int fkmc_comm_destroy(fkmc_ctx* ctx) {
    if (!ctx) return FKMC_ERR_INVALID;
    if (!ctx->nccl_comm) return FKMC_OK;
    nccl_api* api = load_nccl();
    if (api) {
        cudaStreamSynchronize(ctx->stream);
        api->CommDestroy((ncclComm_t)ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    ctx->nccl_nranks = 0;
    return FKMC_OK;
}

// All ranks call this with the same number of chains and measured sweeps.  out_dev (device, 3 * n_measured * nranks * n_chains
// doubles) receives energies | d2energies | c_energies, each as [n_measured][nranks * n_chains] in global chain order; the host
// pointers (any may be NULL) receive copies.  Without a communicator (single GPU) the local series are returned as they are.
extern "C" int fkmc_gather_series(fkmc_ctx* ctx, int* n_measured, int* total_chains, double* energies, double* d2energies, double* c_energies,
                                  void** out_dev) {
    if (!ctx) return FKMC_ERR_INVALID;
    fkmc_chain_state& S = ctx->chain;
    if (!S.active || !S.p.measure_energy) return fkmc_set_error(ctx, FKMC_ERR_STATE, "no energy series to gather");
    FKMC_CUDA(ctx, cudaSetDevice(ctx->device));
    const int nranks = ctx->nccl_comm ? ctx->nccl_nranks : 1;
    const size_t C = S.n_chains, M = (size_t)S.measured, loc = M * C, per = loc * nranks;
    if (n_measured) *n_measured = (int)M;
    if (total_chains) *total_chains = (int)(C * nranks);
    if (per == 0) return FKMC_OK;
    if (3 * per > ctx->gather_cap) {
        if (ctx->d_gather) { FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_gather); ctx->d_gather = nullptr; ctx->gather_cap = 0; }
        FKMC_CUDA(ctx, cudaMalloc(&ctx->d_gather, sizeof(double) * 6 * per));  // gathered + reordered
        ctx->gather_cap = 3 * per;
    }
    double* raw = ctx->d_gather;
    double* ord = ctx->d_gather + ctx->gather_cap;
    const double* src[3] = {S.s_energy, S.s_d2energy, S.s_cenergy};
    if (nranks == 1) {
        for (int o = 0; o < 3; ++o) FKMC_CUDA(ctx, cudaMemcpyAsync(ord + o * per, src[o], sizeof(double) * loc, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        nccl_api* api = load_nccl();
        if (!api) return fkmc_set_error(ctx, FKMC_ERR_STATE, nccl_load_error());
        {
            fkmc_prof_scope ps(ctx, "gather");
            for (int o = 0; o < 3; ++o)
                FKMC_NCCL(ctx, api, api->AllGather(src[o], raw + o * per, loc, ncclDouble, (ncclComm_t)ctx->nccl_comm, ctx->stream));
            reorder_series_kernel<<<dim3(148, 3), 256, 0, ctx->stream>>>(raw, ord, nranks, (int)M, (int)C);
            ctx->launches++;
        }
        FKMC_CUDA(ctx, cudaGetLastError());
    }
    double* host[3] = {energies, d2energies, c_energies};
    for (int o = 0; o < 3; ++o)
        if (host[o]) FKMC_CUDA(ctx, cudaMemcpyAsync(host[o], ord + o * per, sizeof(double) * per, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_dev) *out_dev = ord;
    FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FKMC_OK;
}
