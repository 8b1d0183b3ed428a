// Lattice stencil tables (host) and dense-Hamiltonian assembly (device).
//
// Replaces hypercubic_lattice<D>::{index_to_pos,pos_to_index} (src/lattice/hypercubic.cpp:31-51),
// fill_nearest_neighbors (:116-131), fill_triangular (:137-155), fill_honeycomb (:160-203) and
// configuration_t::calc_hamiltonian (src/configuration.cpp:79-91) + the sparse->dense copy at :212.
// The stencil is kept as a slot-major neighbour table nbr[z][site] so that device code reads it
// coalesced; slot z carries its own hopping value per site, padding slots point at the "zero slot" N.
#include <algorithm>

#include "common.cuh"

namespace {

struct site_pos {
    int c[3];
};

inline site_pos to_pos(int index, int ndim, int L) {
    site_pos p{{0, 0, 0}};
    for (int i = ndim - 1; i >= 0; --i) {
        p.c[i] = index % L;
        index /= L;
    }
    return p;
}
inline int to_index(const site_pos& p, int ndim, int L) {
    int out = 0;
    for (int i = 0; i < ndim; ++i) out = out * L + p.c[i];
    return out;
}
inline int wrap(int x, int L) { return (x % L + L) % L; }

}  // namespace

int fkmc_build_lattice(fkmc_ctx* ctx) {
    const int L = ctx->L;
    switch (ctx->kind) {
        case FKMC_CUBIC1D: ctx->ndim = 1; break;
        case FKMC_CUBIC2D: case FKMC_TRIANGULAR: case FKMC_HONEYCOMB: case FKMC_HONEYCOMB_REF_LOWER: ctx->ndim = 2; break;
        case FKMC_CUBIC3D: ctx->ndim = 3; break;
        default: return fkmc_set_error(ctx, FKMC_ERR_INVALID, "unknown lattice kind");
    }
    if (L < 3) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "L >= 3 required (reference inserts duplicate hoppings for L < 3)");
    if ((ctx->kind == FKMC_HONEYCOMB || ctx->kind == FKMC_HONEYCOMB_REF_LOWER) && L % 2 != 0)
        return fkmc_set_error(ctx, FKMC_ERR_INVALID, "Need even size");  // hypercubic.cpp:182
    int N = 1;
    for (int d = 0; d < ctx->ndim; ++d) N *= L;
    ctx->N = N;
    const int nd = ctx->ndim;
    std::vector<std::vector<std::pair<int, double>>> adj(N);
    auto bond = [&](int i, const site_pos& q, double v) { adj[i].push_back({to_index(q, nd, L), v}); };
    for (int i = 0; i < N; ++i) {
        const site_pos p = to_pos(i, nd, L);
        if (ctx->kind == FKMC_CUBIC1D || ctx->kind == FKMC_CUBIC2D || ctx->kind == FKMC_CUBIC3D || ctx->kind == FKMC_TRIANGULAR) {
            for (int d = 0; d < nd; ++d) {
                site_pos a = p, b = p;
                a.c[d] = wrap(p.c[d] - 1, L);
                b.c[d] = wrap(p.c[d] + 1, L);
                bond(i, a, -ctx->t);
                bond(i, b, -ctx->t);
            }
            if (ctx->kind == FKMC_TRIANGULAR) {
                site_pos a = p, b = p;
                for (int d = 0; d < 2; ++d) {
                    a.c[d] = wrap(p.c[d] - 1, L);
                    b.c[d] = wrap(p.c[d] + 1, L);
                }
                bond(i, a, -ctx->tp);
                bond(i, b, -ctx->tp);
            }
        } else {
            // brick wall: x = last coordinate, y = first; horizontal bonds everywhere, one vertical bond
            site_pos l = p, r = p, u = p, d = p;
            l.c[1] = wrap(p.c[1] - 1, L);
            r.c[1] = wrap(p.c[1] + 1, L);
            d.c[0] = wrap(p.c[0] - 1, L);
            u.c[0] = wrap(p.c[0] + 1, L);
            const bool subA = (ctx->kind == FKMC_HONEYCOMB) ? ((p.c[0] + p.c[1]) % 2 == 0) : (i % 2 == 0);
            bond(i, l, -ctx->t);
            bond(i, r, -ctx->t);
            bond(i, subA ? u : d, -ctx->t);
        }
    }
    if (ctx->kind == FKMC_HONEYCOMB_REF_LOWER) {
        // what a lower-triangle dense solver sees of the literal (non-symmetric) matrix: entries
        // (row i -> col j) with i > j, mirrored
        std::vector<std::vector<std::pair<int, double>>> sym(N);
        for (int i = 0; i < N; ++i)
            for (auto& e : adj[i])
                if (i > e.first) {
                    sym[i].push_back(e);
                    sym[e.first].push_back({i, e.second});
                }
        adj.swap(sym);
    }
    size_t Z = 0;
    for (auto& a : adj) Z = std::max(Z, a.size());
    if (Z > FKMC_MAX_Z) return fkmc_set_error(ctx, FKMC_ERR_INVALID, "too many neighbours");
    ctx->Z = (int)Z;
    ctx->h_nbr_idx.assign(Z * N, N);
    ctx->h_nbr_val.assign(Z * N, 0.0);
    for (int i = 0; i < N; ++i)
        for (size_t z = 0; z < adj[i].size(); ++z) {
            ctx->h_nbr_idx[z * N + i] = adj[i][z].first;
            ctx->h_nbr_val[z * N + i] = adj[i][z].second;
        }
    return FKMC_OK;
}

// One CTA per (32-column block, matrix): zero the lower part of the block's columns, then
// scatter the stencil entries and the diagonal.
__global__ void __launch_bounds__(256) build_h_kernel(const int32_t* __restrict__ f, const int* __restrict__ nbr_idx,
                                                      const double* __restrict__ nbr_val, int N, int Z, double U, double mu_c,
                                                      double* __restrict__ A_all) {
    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    double* A = A_all + (size_t)b * N * N;
    const int32_t* fb = f + (size_t)b * N;
    const int rows = N - c0;
    const int ncols = min(32, N - c0);
    for (int idx = threadIdx.x; idx < rows * ncols; idx += blockDim.x) {
        const int c = c0 + idx / rows, r = c0 + idx % rows;
        A[(size_t)c * N + r] = 0.0;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < ncols * (Z + 1); idx += blockDim.x) {
        const int c = c0 + idx % ncols, z = idx / ncols;
        if (z == Z) {
            A[(size_t)c * N + c] = U * (double)fb[c] - mu_c;
        } else {
            const int r = nbr_idx[z * N + c];
            if (r < N && r > c) A[(size_t)c * N + r] = nbr_val[z * N + c];
        }
    }
}

int fkmc_launch_build_h(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double* d_A) {
    fkmc_prof_scope ps(ctx, "build_h");
    dim3 grid((ctx->N + 31) / 32, B);
    build_h_kernel<<<grid, 256, 0, ctx->stream>>>(d_f, ctx->d_nbr_idx, ctx->d_nbr_val, ctx->N, ctx->Z, U, mu_c, d_A);
    ctx->launches++;
    FKMC_CUDA(ctx, cudaGetLastError());
    return FKMC_OK;
}
