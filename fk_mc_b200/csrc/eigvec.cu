// Eigenvector path, used on measurement sweeps only (IPR, eigenfunctions, stiffness):
// configuration_t::calc_ed(true) (src/configuration.cpp:213,216-219) and measure_ipr::accumulate
// (include/fk_mc/measures/ipr.hpp:39-56).
//
//   H --(one-stage blocked sytrd, reflectors kept in A)--> T --(Sturm bisection)--> eigenvalues
//     --(inverse iteration on T, tridiagonal LU with partial pivoting, one thread per eigenvalue,
//        modified Gram-Schmidt inside clusters of close eigenvalues)--> eigenvectors of T
//     --(back-transformation Z = H_0 H_1 ... H_{n-2} Z_T, 16 eigenvector columns per CTA held in shared
//        memory, IPR fused into the epilogue)--> eigenvectors of H, column k <-> eigenvalue k.
// The inverse iteration follows the structure of LAPACK dstein/dlagtf/dlagts (restated, not copied): it is
// what the reference's Eigen solver delivers up to the choice of basis inside degenerate subspaces.
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "common.cuh"

namespace {

constexpr int BT_COLS = 16;   // eigenvector columns per back-transformation CTA
constexpr int BT_THREADS = 256;

// scratch layout: element i of eigenvalue k of array q lives at ((q * N + i) * N + k)  (coalesced across k)
struct stein_args {
    const double* d;       // [B][N]
    const double* e;       // [B][N]
    const double* evals;   // [B][N] ascending
    double* zt;            // [B][N][N]  zt[i*N + k] = component i of tridiagonal eigenvector k
    double* scratch;       // [B][5][N][N]
    int N;
};

__device__ __forceinline__ double hash_unit(unsigned k, unsigned i, unsigned it) {
    unsigned h = k * 2654435761u ^ (i + 0x9e3779b9u) * 2246822519u ^ (it * 3266489917u);
    h ^= h >> 15; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return (double)h * (2.0 / 4294967296.0) - 1.0;
}

// One thread per eigenvalue; the first eigenvalue of every cluster (gap to its predecessor > ortol) computes the whole
// cluster sequentially so that the Gram-Schmidt sweep can see the finished predecessors.
__global__ void __launch_bounds__(128) stein_kernel(stein_args P) {
    const int N = P.N, b = blockIdx.y;
    const int k0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (k0 >= N) return;
    const double* d = P.d + (size_t)b * N;
    const double* e = P.e + (size_t)b * N;
    const double* w = P.evals + (size_t)b * N;
    double* zt = P.zt + (size_t)b * N * N;
    double* S = P.scratch + (size_t)b * 5 * N * N;
    const size_t NN = (size_t)N * N;
    double* fa = S + k0;            // pivots           fa[i*N]
    double* fb = S + NN + k0;       // first super-diagonal of U
    double* fc = S + 2 * NN + k0;   // multipliers
    double* fd = S + 3 * NN + k0;   // second super-diagonal of U
    double* fi = S + 4 * NN + k0;   // interchange flags (0 / 1)

    // 1-norm of T
    double onenrm = fabs(d[0]) + (N > 1 ? fabs(e[0]) : 0.0);
    for (int i = 1; i < N; ++i) onenrm = fmax(onenrm, fabs(d[i]) + fabs(e[i - 1]) + (i < N - 1 ? fabs(e[i]) : 0.0));
    if (onenrm == 0.0) onenrm = 1.0;
    const double eps = DBL_EPSILON;
    const double ortol = 1e-6 * onenrm;
    if (k0 > 0 && w[k0] - w[k0 - 1] <= ortol) return;  // not the first of its cluster

    double xjm = 0.0;
    for (int j = k0; j < N && (j == k0 || w[j] - w[j - 1] <= ortol); ++j) {
        double xj = w[j];
        if (j > k0) {  // keep the shifts of a cluster distinct (dstein: pertol = 10 eps |xj|)
            const double pertol = 10.0 * fabs(eps * xj) + 10.0 * eps * eps * onenrm;
            if (xj - xjm < pertol) xj = xjm + pertol;
        }
        xjm = xj;
        double* z = zt + j;  // z[i*N]
        // ---- LU factorisation of T - xj I with partial pivoting (row interchanges of adjacent rows) ----
        const double tol = eps * onenrm;
        double ak = d[0] - xj;                 // current pivot row: (ak, bk) on columns k, k+1
        double bk = (N > 1) ? e[0] : 0.0;
        for (int k = 0; k < N - 1; ++k) {
            const double ck = e[k];            // sub-diagonal entry of row k+1 in column k
            const double a1 = d[k + 1] - xj;   // row k+1: (ck, a1, b1)
            const double b1 = (k < N - 2) ? e[k + 1] : 0.0;
            if (fabs(ck) <= fabs(ak)) {        // no interchange
                const double mult = (ak != 0.0) ? ck / ak : 0.0;
                fa[(size_t)k * N] = ak;
                fb[(size_t)k * N] = bk;
                fd[(size_t)k * N] = 0.0;
                fc[(size_t)k * N] = mult;
                fi[(size_t)k * N] = 0.0;
                ak = a1 - mult * bk;
                bk = b1;
            } else {                           // interchange rows k and k+1
                const double mult = ak / ck;
                fa[(size_t)k * N] = ck;
                fb[(size_t)k * N] = a1;
                fd[(size_t)k * N] = b1;
                fc[(size_t)k * N] = mult;
                fi[(size_t)k * N] = 1.0;
                ak = bk - mult * a1;
                bk = -mult * b1;
            }
        }
        fa[(size_t)(N - 1) * N] = ak;
        fb[(size_t)(N - 1) * N] = 0.0;
        fd[(size_t)(N - 1) * N] = 0.0;
        // ---- inverse iteration ----
        for (int i = 0; i < N; ++i) z[(size_t)i * N] = hash_unit((unsigned)j, (unsigned)i, 0u);
        int good = 0;
        for (int its = 0; its < 6 && good < 2; ++its) {
            // scale the right-hand side (dstein: n * onenrm * max(eps, |last pivot|) / ||x||_1)
            double n1 = 0.0;
            for (int i = 0; i < N; ++i) n1 += fabs(z[(size_t)i * N]);
            const double scl = (double)N * onenrm * fmax(eps, fabs(fa[(size_t)(N - 1) * N])) / fmax(n1, DBL_MIN);
            for (int i = 0; i < N; ++i) z[(size_t)i * N] *= scl;
            // forward: apply the interchanges and multipliers
            for (int k = 0; k < N - 1; ++k) {
                const double yk = z[(size_t)k * N], yk1 = z[(size_t)(k + 1) * N], m = fc[(size_t)k * N];
                if (fi[(size_t)k * N] == 0.0) {
                    z[(size_t)(k + 1) * N] = yk1 - m * yk;
                } else {
                    z[(size_t)k * N] = yk1;
                    z[(size_t)(k + 1) * N] = yk - m * yk1;
                }
            }
            // backward: U x = y with tiny pivots replaced by +-tol
            double x1 = 0.0, x2 = 0.0;
            for (int k = N - 1; k >= 0; --k) {
                double piv = fa[(size_t)k * N];
                if (fabs(piv) < tol) piv = (piv < 0.0) ? -tol : tol;
                const double t = z[(size_t)k * N] - fb[(size_t)k * N] * x1 - fd[(size_t)k * N] * x2;
                const double x = t / piv;
                z[(size_t)k * N] = x;
                x2 = x1;
                x1 = x;
            }
            // re-orthogonalise against the finished vectors of this cluster (modified Gram-Schmidt)
            for (int q = k0; q < j; ++q) {
                const double* zq = zt + q;
                double dot = 0.0;
                for (int i = 0; i < N; ++i) dot = fma(z[(size_t)i * N], zq[(size_t)i * N], dot);
                for (int i = 0; i < N; ++i) z[(size_t)i * N] = fma(-dot, zq[(size_t)i * N], z[(size_t)i * N]);
            }
            // growth check + normalisation
            double nrm2 = 0.0, amax = 0.0;
            for (int i = 0; i < N; ++i) {
                const double x = z[(size_t)i * N];
                nrm2 = fma(x, x, nrm2);
                amax = fmax(amax, fabs(x));
            }
            if (!(nrm2 > 0.0) || !isfinite(nrm2)) {  // breakdown: restart from another pseudo-random vector
                for (int i = 0; i < N; ++i) z[(size_t)i * N] = hash_unit((unsigned)j, (unsigned)i, (unsigned)its + 1u);
                continue;
            }
            const double inv = 1.0 / sqrt(nrm2);
            for (int i = 0; i < N; ++i) z[(size_t)i * N] *= inv;
            if (amax >= sqrt(0.1 / (double)N)) ++good;  // the solve amplified the vector enough: count a converged step
        }
    }
}

// Z = H_0 ... H_{n-2} Z_T for a block of 16 eigenvector columns, then IPR.  A holds the reflectors of the one-stage
// sytrd (v_i in column i below the diagonal, leading 1 explicit at row i+1), tau[i] their scalars.
__global__ void __launch_bounds__(BT_THREADS, 1)
backtransform_kernel(const double* __restrict__ A_all, const double* __restrict__ tau_all, const double* __restrict__ zt_all, int N,
                     double* __restrict__ evecs_all /*[B][N][N] col-major or null*/, double* __restrict__ ipr_all /*[B][N] or null*/,
                     double* __restrict__ vt_all /*[B][N][N] site-major (vt[i][k]) or null*/) {
    extern __shared__ double zb[];  // [N][17] (padded rows)
    __shared__ double wpart[BT_THREADS / 32][BT_COLS];
    __shared__ double wfull[BT_COLS];
    const int LDZ = BT_COLS + 1;
    const int b = blockIdx.y, c0 = blockIdx.x * BT_COLS, tid = threadIdx.x;
    const int c = tid & (BT_COLS - 1), rl = tid >> 4;  // column inside the block, row lane (0..15)
    const int lane = tid & 31, warp = tid >> 5;
    const double* A = A_all + (size_t)b * N * N;
    const double* tau = tau_all + (size_t)b * N;
    const double* zt = zt_all + (size_t)b * N * N;
    const int ncol = min(BT_COLS, N - c0);
    for (int idx = tid; idx < N * BT_COLS; idx += BT_THREADS) {
        const int i = idx / BT_COLS, cc = idx % BT_COLS;
        zb[i * LDZ + cc] = (cc < ncol) ? zt[(size_t)i * N + c0 + cc] : 0.0;
    }
    __syncthreads();
    for (int i = N - 2; i >= 0; --i) {
        const double t = tau[i];
        if (t == 0.0) continue;  // uniform
        const double* v = A + (size_t)i * N;  // v[r], r > i; v[i+1] = 1
        // row ownership must not depend on i: thread (c, rl) owns the rows r == rl (mod 16) of column c for the whole kernel
        const int rs = i + 1 + ((rl - (i + 1)) & 15);
        double part = 0.0;
        for (int r = rs; r < N; r += 16) part = fma(v[r], zb[r * LDZ + c], part);
        part += __shfl_xor_sync(0xffffffffu, part, 16);  // the two row lanes of this warp
        if (lane < 16) wpart[warp][c] = part;
        __syncthreads();
        if (tid < BT_COLS) {
            double s = 0.0;
#pragma unroll
            for (int ww = 0; ww < BT_THREADS / 32; ++ww) s += wpart[ww][tid];
            wfull[tid] = t * s;
        }
        __syncthreads();
        const double wc = wfull[c];
        for (int r = rs; r < N; r += 16) zb[r * LDZ + c] = fma(-v[r], wc, zb[r * LDZ + c]);
        // (element (r, c) is only ever touched by thread (c, r mod 16), so no barrier is needed before the next reflector)
    }
    __syncthreads();
    if (evecs_all) {
        double* ev = evecs_all + (size_t)b * N * N;
        for (int idx = tid; idx < N * BT_COLS; idx += BT_THREADS) {
            const int cc = idx / N, i = idx % N;
            if (cc < ncol) ev[(size_t)(c0 + cc) * N + i] = zb[i * LDZ + cc];
        }
    }
    if (vt_all) {
        double* vt = vt_all + (size_t)b * N * N;
        for (int idx = tid; idx < N * BT_COLS; idx += BT_THREADS) {
            const int i = idx / BT_COLS, cc = idx % BT_COLS;
            if (cc < ncol) vt[(size_t)i * N + c0 + cc] = zb[i * LDZ + cc];
        }
    }
    if (ipr_all) {
        // ipr_k = ||psi||_4 / ||psi||_2^2  (include/fk_mc/measures/ipr.hpp:47-53)
        double s2 = 0.0, s4 = 0.0;
        for (int r = rl; r < N; r += 16) {
            const double x = zb[r * LDZ + c], x2 = x * x;
            s2 += x2;
            s4 = fma(x2, x2, s4);
        }
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
        s4 += __shfl_xor_sync(0xffffffffu, s4, 16);
        __shared__ double p2[BT_THREADS / 32][BT_COLS], p4[BT_THREADS / 32][BT_COLS];
        if (lane < 16) { p2[warp][c] = s2; p4[warp][c] = s4; }
        __syncthreads();
        if (tid < ncol) {
            double a2 = 0.0, a4 = 0.0;
            for (int ww = 0; ww < BT_THREADS / 32; ++ww) { a2 += p2[ww][tid]; a4 += p4[ww][tid]; }
            ipr_all[(size_t)b * N + c0 + tid] = sqrt(sqrt(a4)) / a2;
        }
    }
}


// ---- blocked back-transformation on the FP64 tensor cores ----------------------------------------------------------------------
// The reflectors of the one-stage tridiagonalisation are grouped 32 at a time into compact-WY block reflectors,
// H_{i0} ... H_{i0+31} = I - V T V^T  (T upper triangular, LAPACK dlarft "forward, columnwise"), and applied from the last group to
// the first:  Z <- Z - V (T (V^T Z)).  Both products are DMMA GEMMs; a CTA keeps 16 eigenvector columns in shared memory for the
// whole kernel and streams V from L2.  Three block barriers per GROUP instead of two per REFLECTOR.
constexpr int BG = 32;          // reflectors per group
constexpr int ZLD = 20;         // row stride of the Z block in shared memory (== 4 mod 16: conflict-free B-operand reads)

// T factors: grid (ngroups, B), 256 threads.  V[r][q] = A[(i0+q) N + r] for r > i0+q (leading 1 stored explicitly), 0 above.
__global__ void __launch_bounds__(256) larft_kernel(const double* __restrict__ A_all, const double* __restrict__ tau_all, int N, int ngroups,
                                                    double* __restrict__ T_all) {
    __shared__ double Vs[32][33];
    __shared__ double G[32][33];
    const int grp = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int i0 = grp * BG;
    const double* A = A_all + (size_t)b * N * N;
    const double* tau = tau_all + (size_t)b * N;
    // this thread's 4 (a, c) pairs of the Gram matrix: a = w + 8 u, c = lane
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r0 = i0 + 1; r0 < N; r0 += 32) {
        // tile Vs[rr][q], rr = row r0 + rr: warp w loads columns q = w, w+8, w+16, w+24 (coalesced along r)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = w + 8 * u, r = r0 + lane, col = i0 + q;
            Vs[lane][q] = (r < N && col < N - 1 && r > col) ? A[(size_t)col * N + r] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = w + 8 * u;
            double sacc = acc[u];
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) sacc = fma(Vs[rr][a], Vs[rr][lane], sacc);
            acc[u] = sacc;
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) G[w + 8 * u][lane] = acc[u];
    __syncthreads();
    if (w == 0) {
        // lane i owns row i of T:  T[i][j] = -tau_j sum_{l=i}^{j-1} T[i][l] G[l][j]  (i < j),  T[j][j] = tau_j
        double tr[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int col = i0 + j;
            const double tj = (col < N - 1) ? tau[col] : 0.0;
            double sacc = 0.0;
#pragma unroll
            for (int l = 0; l < 32; ++l)
                if (l < j) sacc = (l >= lane) ? fma(tr[l], G[l][j], sacc) : sacc;
            tr[j] = (j == lane) ? tj : ((j > lane) ? -tj * sacc : 0.0);
        }
        double* T = T_all + ((size_t)b * ngroups + grp) * BG * BG;
#pragma unroll
        for (int j = 0; j < 32; ++j) T[lane * BG + j] = tr[j];
    }
}

// Z = H_0 ... H_{n-2} Z_T for 16 eigenvector columns per CTA, then the outputs of the per-reflector kernel above.
__global__ void __launch_bounds__(256, 1)
backtransform_wy_kernel(const double* __restrict__ A_all, const double* __restrict__ T_all, const double* __restrict__ zt_all, int N, int ngroups,
                        double* __restrict__ evecs_all, double* __restrict__ ipr_all, double* __restrict__ vt_all) {
    extern __shared__ __align__(16) double zsm[];
    const int Np = (N + 7) & ~7;                 // rows padded to the DMMA row block
    double* Z = zsm;                             // [Np][ZLD]
    double* Wp = Z + (size_t)Np * ZLD;           // [8 warps][32][16] partial V^T Z
    double* W2 = Wp + 8 * 512;                   // [32][17] T (V^T Z)
    double* Ts = W2 + 32 * 17;                   // [32][33]
    const int b = blockIdx.y, c0 = blockIdx.x * 16, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const double* A = A_all + (size_t)b * N * N;
    const double* zt = zt_all + (size_t)b * N * N;
    const int ncol = min(16, N - c0);
    for (int idx = tid; idx < Np * 16; idx += 256) {
        const int i = idx >> 4, cc = idx & 15;
        Z[i * ZLD + cc] = (i < N && cc < ncol) ? zt[(size_t)i * N + c0 + cc] : 0.0;
    }
    __syncthreads();
    for (int grp = ngroups - 1; grp >= 0; --grp) {
        const int i0 = grp * BG;
        const int rlo = (i0 + 1) & ~3;           // first k4 step that can hold a non-zero of V (rows > i0)
        // ---- (1) partial W = V^T Z over this warp's k4 steps: 4 (q blocks) x 2 (column blocks) accumulators ----
        double acc[4][2][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int n = 0; n < 2; ++n) acc[a][n][0] = acc[a][n][1] = 0.0;
        for (int r0 = rlo + 4 * warp; r0 < N; r0 += 32) {
            const int r = r0 + t;
            double af[4], bf[2];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int col = i0 + 8 * a + g;   // A operand: V^T[q = 8a + g][r]
                af[a] = (r < N && col < N - 1 && r > col) ? __ldg(A + (size_t)col * N + r) : 0.0;
            }
#pragma unroll
            for (int n = 0; n < 2; ++n) bf[n] = (r < Np) ? Z[r * ZLD + 8 * n + g] : 0.0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int n = 0; n < 2; ++n) dmma884(acc[a][n][0], acc[a][n][1], af[a], bf[n]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                Wp[warp * 512 + (8 * a + g) * 16 + 8 * n + 2 * t] = acc[a][n][0];
                Wp[warp * 512 + (8 * a + g) * 16 + 8 * n + 2 * t + 1] = acc[a][n][1];
            }
        // T of this group
        const double* T = T_all + ((size_t)b * ngroups + grp) * BG * BG;
        for (int idx = tid; idx < BG * BG; idx += 256) Ts[(idx >> 5) * 33 + (idx & 31)] = T[idx];
        __syncthreads();
        // ---- (2) W = sum of the partials (fixed order); W2 = T W ----
        {
            const int q = tid >> 4, cc = tid & 15;   // 256 threads: rows q and q + 16
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double s = 0.0;
#pragma unroll
                for (int ww = 0; ww < 8; ++ww) s += Wp[ww * 512 + (q + 16 * h) * 16 + cc];
                W2[(q + 16 * h) * 17 + cc] = s;   // W for now
            }
        }
        __syncthreads();
        double w2a = 0.0, w2b = 0.0;
        {
            const int q = tid >> 4, cc = tid & 15;
            for (int l = q; l < BG; ++l) w2a = fma(Ts[q * 33 + l], W2[l * 17 + cc], w2a);           // T upper triangular
            for (int l = q + 16; l < BG; ++l) w2b = fma(Ts[(q + 16) * 33 + l], W2[l * 17 + cc], w2b);
        }
        __syncthreads();
        {
            const int q = tid >> 4, cc = tid & 15;
            W2[q * 17 + cc] = w2a;
            W2[(q + 16) * 17 + cc] = w2b;
        }
        __syncthreads();
        // ---- (3) Z -= V W2 over this warp's 8-row blocks ----
        for (int rb = ((i0 + 1) & ~7) + 8 * warp; rb < N; rb += 64) {
            const int r = rb + g;
            double c[2][2];
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                const double2 v = *reinterpret_cast<const double2*>(Z + r * ZLD + 8 * n + 2 * t);
                c[n][0] = v.x;
                c[n][1] = v.y;
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const int col = i0 + 4 * kk + t;     // A operand: -V[r][q = 4 kk + t]
                const double av = (r < N && col < N - 1 && r > col) ? -__ldg(A + (size_t)col * N + r) : 0.0;
#pragma unroll
                for (int n = 0; n < 2; ++n) dmma884(c[n][0], c[n][1], av, W2[(4 * kk + t) * 17 + 8 * n + g]);
            }
#pragma unroll
            for (int n = 0; n < 2; ++n) *reinterpret_cast<double2*>(Z + r * ZLD + 8 * n + 2 * t) = make_double2(c[n][0], c[n][1]);
        }
        __syncthreads();
    }
    if (evecs_all) {
        double* ev = evecs_all + (size_t)b * N * N;
        for (int idx = tid; idx < N * 16; idx += 256) {
            const int cc = idx / N, i = idx % N;
            if (cc < ncol) ev[(size_t)(c0 + cc) * N + i] = Z[i * ZLD + cc];
        }
    }
    if (vt_all) {
        double* vt = vt_all + (size_t)b * N * N;
        for (int idx = tid; idx < N * 16; idx += 256) {
            const int i = idx >> 4, cc = idx & 15;
            if (cc < ncol) vt[(size_t)i * N + c0 + cc] = Z[i * ZLD + cc];
        }
    }
    if (ipr_all) {
        // ipr_k = ||psi||_4 / ||psi||_2^2  (include/fk_mc/measures/ipr.hpp:47-53): thread (cc, rl) sums rows rl, rl + 16, ...
        const int cc = tid & 15, rl = tid >> 4;
        double s2 = 0.0, s4 = 0.0;
        for (int r = rl; r < N; r += 16) {
            const double x = Z[r * ZLD + cc], x2 = x * x;
            s2 += x2;
            s4 = fma(x2, x2, s4);
        }
        __syncthreads();
        Wp[tid] = s2;
        Wp[256 + tid] = s4;
        __syncthreads();
        if (tid < ncol) {
            double a2 = 0.0, a4 = 0.0;
            for (int rr = 0; rr < 16; ++rr) { a2 += Wp[rr * 16 + tid]; a4 += Wp[256 + rr * 16 + tid]; }
            ipr_all[(size_t)b * N + c0 + tid] = sqrt(sqrt(a4)) / a2;
        }
    }
}

}  // namespace

// Eigen-decomposition of the Hamiltonians of B configurations (device f).  Outputs on the device: evals [B][N],
// optionally evecs [B][N][N] (column-major) and ipr [B][N]; logZ etc. in d_out [B][8].  Processes the batch in chunks.
static int eigvec_pipeline_impl(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out,
                                double* h_evecs, double* h_ipr_host, double* d_ipr, double* d_vt, double* d_evecs_dev = nullptr) {
    const int N = ctx->N;
    const size_t NN = (size_t)N * N;
    if (sizeof(double) * (size_t)N * (BT_COLS + 1) + 4096 > ctx->smem_optin)
        return fkmc_set_error(ctx, FKMC_ERR_INVALID, "eigenvectors: N too large for the back-transformation kernel");
    // chunk of the batch: scratch is 6 N^2 doubles per matrix (+ N^2 for a host copy of the eigenvectors); use up to half of the free device
    // memory, and whole waves of CTAs (the tridiagonalisation runs one CTA per matrix)
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)12e9;
    const double per_matrix = (6.0 + (h_evecs ? 1.0 : 0.0)) * NN * 8.0 + 32.0 * N * 8.0;
    int chunk = (int)std::max<double>(1.0, std::min<double>((double)B, 0.5 * ((double)free_b + 8.0 * (double)ctx->ev_scratch_cap) / per_matrix));
    chunk = std::min(chunk, ctx->max_batch);
    if (chunk < B && chunk > ctx->num_sms) chunk -= chunk % ctx->num_sms;
    // the big scratch (inverse-iteration factors, tridiagonal eigenvectors, T factors) is cached in the context: allocating and freeing
    // tens of GB per call costs more than the kernels
    const int ngroups_a = (N - 1 + BG - 1) / BG;
    const size_t need = (size_t)chunk * (6 * NN + (size_t)ngroups_a * BG * BG);
    if (need > ctx->ev_scratch_cap) {
        if (ctx->d_ev_scratch) {
            FKMC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_ev_scratch);
            ctx->d_ev_scratch = nullptr;
            ctx->ev_scratch_cap = 0;
        }
        FKMC_CUDA(ctx, cudaMalloc(&ctx->d_ev_scratch, sizeof(double) * need));
        ctx->ev_scratch_cap = need;
    }
    double *d_zt = ctx->d_ev_scratch, *d_scr = d_zt + NN * chunk, *d_ev_host = nullptr, *d_ip = nullptr;
    if (h_evecs) FKMC_CUDA(ctx, cudaMalloc(&d_ev_host, sizeof(double) * NN * chunk));
    if (!d_ipr && h_ipr_host) FKMC_CUDA(ctx, cudaMalloc(&d_ip, sizeof(double) * (size_t)N * chunk));
    int rc = FKMC_OK;
    const size_t smem = sizeof(double) * (size_t)N * (BT_COLS + 1);
    cudaFuncSetAttribute(backtransform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // blocked (compact-WY, DMMA) back-transformation when its shared-memory footprint fits; "eigvec_v1" = 1 keeps the per-reflector kernel
    const int ngroups = (N - 1 + BG - 1) / BG;
    const size_t smem_wy = sizeof(double) * ((size_t)((N + 7) & ~7) * ZLD + 8 * 512 + 32 * 17 + 32 * 33);
    const bool use_wy = !ctx->eigvec_v1 && N >= 8 && smem_wy <= ctx->smem_optin;
    double* d_T = d_scr + 5 * NN * chunk;
    if (use_wy) {
        cudaFuncSetAttribute(backtransform_wy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_wy);
    }
    for (int b0 = 0; b0 < B && !rc; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        const int32_t* f = d_f + (size_t)b0 * N;
        // one-stage tridiagonalisation keeps the reflectors needed for the back-transformation
        if ((rc = fkmc_launch_build_h(ctx, f, nb, U, mu_c, ctx->d_A))) break;
        if ((rc = fkmc_launch_sytrd(ctx, ctx->d_A, N, nb, ctx->d_d, ctx->d_e, ctx->d_tau, ctx->d_W))) break;
        if ((rc = fkmc_launch_tridiag_eig(ctx, ctx->d_d, ctx->d_e, N, nb, beta, d_evals + (size_t)b0 * N, N, nullptr, 0, d_out + (size_t)b0 * 8,
                                          nullptr, nullptr)))
            break;
        {
            fkmc_prof_scope ps(ctx, "stein");
            stein_args P{ctx->d_d, ctx->d_e, d_evals + (size_t)b0 * N, d_zt, d_scr, N};
            dim3 grid((N + 127) / 128, nb);
            stein_kernel<<<grid, 128, 0, ctx->stream>>>(P);
            ctx->launches++;
        }
        double* iprp = d_ipr ? d_ipr + (size_t)b0 * N : d_ip;
        {
            fkmc_prof_scope ps(ctx, "backtransform");
            dim3 grid((N + BT_COLS - 1) / BT_COLS, nb);
            double* vtp = d_vt ? d_vt + (size_t)b0 * NN : nullptr;
            double* d_ev = d_evecs_dev ? d_evecs_dev + (size_t)b0 * NN : d_ev_host;
            if (use_wy) {
                larft_kernel<<<dim3(ngroups, nb), 256, 0, ctx->stream>>>(ctx->d_A, ctx->d_tau, N, ngroups, d_T);
                backtransform_wy_kernel<<<grid, 256, smem_wy, ctx->stream>>>(ctx->d_A, d_T, d_zt, N, ngroups, d_ev, iprp, vtp);
                ctx->launches += 2;
            } else {
                backtransform_kernel<<<grid, BT_THREADS, smem, ctx->stream>>>(ctx->d_A, ctx->d_tau, d_zt, N, d_ev, iprp, vtp);
                ctx->launches++;
            }
        }
        if (cudaGetLastError() != cudaSuccess) { rc = fkmc_set_error(ctx, FKMC_ERR_CUDA, "eigenvector kernels failed to launch"); break; }
        if (h_evecs) cudaMemcpyAsync(h_evecs + (size_t)b0 * NN, d_ev_host, sizeof(double) * NN * nb, cudaMemcpyDeviceToHost, ctx->stream);
        if (h_ipr_host && !d_ipr) cudaMemcpyAsync(h_ipr_host + (size_t)b0 * N, d_ip, sizeof(double) * (size_t)N * nb, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = fkmc_set_error(ctx, FKMC_ERR_CUDA, cudaGetErrorString(cudaGetLastError())); break; }
    }
    cudaFree(d_ev_host); cudaFree(d_ip);
    return rc;
}

int fkmc_eigvec_pipeline(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out,
                         double* h_evecs, double* h_ipr_host, double* d_ipr) {
    return eigvec_pipeline_impl(ctx, d_f, B, U, mu_c, beta, d_evals, d_out, h_evecs, h_ipr_host, d_ipr, nullptr);
}

int fkmc_eigvec_pipeline_dev(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out, double* d_vt) {
    int rc = fkmc_ensure_dense_ws(ctx);
    if (rc) return rc;
    return eigvec_pipeline_impl(ctx, d_f, B, U, mu_c, beta, d_evals, d_out, nullptr, nullptr, nullptr, d_vt);
}

// both eigenvector layouts on the device: evecs [B][N][N] eigenvector-major (evecs[b][k][i]) and vt [B][N][N] site-major (vt[b][i][k])
int fkmc_eigvec_pipeline_dev2(fkmc_ctx* ctx, const int32_t* d_f, int B, double U, double mu_c, double beta, double* d_evals, double* d_out, double* d_evecs,
                              double* d_vt, double* d_ipr) {
    int rc = fkmc_ensure_dense_ws(ctx);
    if (rc) return rc;
    return eigvec_pipeline_impl(ctx, d_f, B, U, mu_c, beta, d_evals, d_out, nullptr, nullptr, d_ipr, d_vt, d_evecs);
}
