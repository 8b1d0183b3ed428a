// Probe: validates the FP64 mma.sync fragment layouts used by fk_mc_b200/csrc and measures
// the DMMA / DFMA peaks on this GPU (the FP64-tensor roofline denominator is not in
// MEASURED_PEAKS.json).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_dmma probe_dmma.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---------------- layout checks: C(MxN) = A(MxK) * B(KxN), A row-major [M][K], B as [K][N] -------------
template <int SHAPE> __global__ void layout_check(const double *A, const double *B, double *C) {
    int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    if (SHAPE == 0) {          // m8n8k4
        double c0 = 0, c1 = 0;
        mma884(c0, c1, A[g * 4 + t], B[t * 8 + g]);
        C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1;
    } else if (SHAPE == 1) {   // m16n8k4
        double c[4] = {0, 0, 0, 0}; double a[2] = {A[g * 4 + t], A[(g + 8) * 4 + t]};
        mma1684(c, a, B[t * 8 + g]);
        C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1]; C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
    } else if (SHAPE == 2) {   // m16n8k8
        double c[4] = {0, 0, 0, 0};
        double a[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + t + 4], A[(g + 8) * 8 + t + 4]};
        double b[2] = {B[t * 8 + g], B[(t + 4) * 8 + g]};
        mma1688(c, a, b);
        C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1]; C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
    } else {                   // m16n8k16
        double c[4] = {0, 0, 0, 0}; double a[8], b[4];
        for (int i = 0; i < 8; i++) a[i] = A[(g + 8 * (i & 1)) * 16 + t + 4 * (i >> 1)];
        for (int i = 0; i < 4; i++) b[i] = B[(t + 4 * i) * 8 + g];
        mma16816(c, a, b);
        C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1]; C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
    }
}

static void run_layout(int shape, int M, int N, int K, const char *name) {
    std::vector<double> A(M * K), B(K * N), C(M * N), R(M * N, 0.0);
    for (auto &x : A) x = (rand() % 17 - 8) / 4.0;
    for (auto &x : B) x = (rand() % 13 - 6) / 8.0;
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) for (int k = 0; k < K; k++) R[i * N + j] += A[i * K + k] * B[k * N + j];
    double *dA, *dB, *dC; CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, B.size() * 8)); CK(cudaMalloc(&dC, C.size() * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
    if (shape == 0) layout_check<0><<<1, 32>>>(dA, dB, dC); else if (shape == 1) layout_check<1><<<1, 32>>>(dA, dB, dC);
    else if (shape == 2) layout_check<2><<<1, 32>>>(dA, dB, dC); else layout_check<3><<<1, 32>>>(dA, dB, dC);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(C.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost));
    double err = 0; for (int i = 0; i < M * N; i++) err = fmax(err, fabs(C[i] - R[i]));
    printf("layout %-10s max_err %.3e %s\n", name, err, err == 0 ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
}

// ---------------- throughput ----------------
template <int SHAPE, int NACC> __global__ void __launch_bounds__(256) tput(double *out, int iters) {
    double seed = threadIdx.x * 1e-3;
    if (SHAPE == 0) {
        double c[NACC][2]; for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0;
        double a = seed, b = 1.0 - seed;
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < NACC; i++) mma884(c[i][0], c[i][1], a, b);
        double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
        if (s == 123.456) out[0] = s;
    } else if (SHAPE == 2) {
        double c[NACC][4]; for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
        double a[4] = {seed, seed + 1, seed + 2, seed + 3}, b[2] = {1 - seed, 2 - seed};
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < NACC; i++) mma1688(c[i], a, b);
        double s = 0; for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
        if (s == 123.456) out[0] = s;
    } else if (SHAPE == 3) {
        double c[NACC][4]; for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
        double a[8], b[4]; for (int j = 0; j < 8; j++) a[j] = seed + j; for (int j = 0; j < 4; j++) b[j] = j - seed;
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < NACC; i++) mma16816(c[i], a, b);
        double s = 0; for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
        if (s == 123.456) out[0] = s;
    } else {  // DFMA
        double c[NACC]; for (int i = 0; i < NACC; i++) c[i] = seed + i;
        double a = 1.0000001, b = seed * 1e-9;
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
        double s = 0; for (int i = 0; i < NACC; i++) s += c[i];
        if (s == 123.456) out[0] = s;
    }
}

template <int SHAPE, int NACC> static void run_tput(const char *name, double flop_per_inst_per_warp, int ctas_per_sm, int nsm) {
    double *out; CK(cudaMalloc(&out, 8));
    int iters = 20000; int grid = nsm * ctas_per_sm;
    tput<SHAPE, NACC><<<grid, 256>>>(out, 100); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); tput<SHAPE, NACC><<<grid, 256>>>(out, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warps = double(grid) * 8; double flops = warps * iters * NACC * flop_per_inst_per_warp;
    printf("tput %-10s acc=%d ctas/sm=%d : %.3f ms  %.2f TFLOP/s\n", name, NACC, ctas_per_sm, best, flops / best * 1e-9);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0, l2 = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0);
    printf("device %s sm=%d cc=%d.%d smem_optin=%zu l2=%d MB clock=%d kHz\n", p.name, p.multiProcessorCount, p.major, p.minor,
           p.sharedMemPerBlockOptin, l2 >> 20, clk);
    run_layout(0, 8, 8, 4, "m8n8k4"); run_layout(1, 16, 8, 4, "m16n8k4"); run_layout(2, 16, 8, 8, "m16n8k8"); run_layout(3, 16, 8, 16, "m16n8k16");
    int nsm = p.multiProcessorCount;
    run_tput<0, 8>("m8n8k4", 2.0 * 8 * 8 * 4, 4, nsm);
    run_tput<0, 16>("m8n8k4", 2.0 * 8 * 8 * 4, 8, nsm);
    run_tput<2, 8>("m16n8k8", 2.0 * 16 * 8 * 8, 4, nsm);
    run_tput<2, 8>("m16n8k8", 2.0 * 16 * 8 * 8, 8, nsm);
    run_tput<3, 8>("m16n8k16", 2.0 * 16 * 8 * 16, 4, nsm);
    run_tput<3, 8>("m16n8k16", 2.0 * 16 * 8 * 16, 8, nsm);
    run_tput<9, 16>("dfma", 2.0 * 32, 8, nsm);
    run_tput<9, 8>("dfma", 2.0 * 32, 4, nsm);
    return 0;
}
