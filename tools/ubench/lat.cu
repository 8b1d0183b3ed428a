// Dependent-issue latencies on sm_100a: DFMA, DADD, 64-bit shuffle, DMMA m8n8k4 accumulate chain, rsqrt/rcp seeds.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* t, double x0) {
    const int N = 256;
    double x = x0 + threadIdx.x, y = 1.0000001;
    long long a, b;
    a = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, 0.5);
    b = clock64(); if (threadIdx.x == 0) t[0] = (b - a);
    a = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + y;
    b = clock64(); if (threadIdx.x == 0) t[1] = (b - a);
    a = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));
    b = clock64(); if (threadIdx.x == 0) t[2] = (b - a);
    double c0 = x, c1 = y;
    a = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) dmma(c0, c1, y, y);
    b = clock64(); if (threadIdx.x == 0) t[3] = (b - a);
    x += c0 + c1;
    a = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    b = clock64(); if (threadIdx.x == 0) t[4] = (b - a);
    a = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
    b = clock64(); if (threadIdx.x == 0) t[5] = (b - a);
    a = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = sqrt(x + 2.0);
    b = clock64(); if (threadIdx.x == 0) t[6] = (b - a);
    a = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = 1.0 / (x + 2.0);
    b = clock64(); if (threadIdx.x == 0) t[7] = (b - a);
    // independent DMMAs (throughput, one warp): 8 accumulators
    double d[16];
    for (int i = 0; i < 16; ++i) d[i] = x + i;
    a = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma(d[2 * j], d[2 * j + 1], y, y);
    }
    b = clock64(); if (threadIdx.x == 0) t[8] = (b - a);
    for (int i = 0; i < 16; ++i) x += d[i];
    // shuffle + add stage (one butterfly stage)
    a = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x += __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15));
    b = clock64(); if (threadIdx.x == 0) t[9] = (b - a);
    out[threadIdx.x] = x;
}
int main() {
    double* o; long long* t; cudaMalloc(&o, 8 * 1024); cudaMalloc(&t, 8 * 16);
    for (int nw = 1; nw <= 8; nw *= 2) {
        k<<<1, 32 * nw>>>(o, t, 1.5); cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, t, 8 * 16, cudaMemcpyDeviceToHost);
        const char* nm[] = {"DFMA", "DADD", "SHFL64", "DMMA chain", "rsqrt.approx", "rcp.approx", "sqrt(x+2)", "1/(x+2)", "DMMA 8 indep", "shfl+add"};
        printf("warps %d:", nw);
        for (int i = 0; i < 10; ++i) printf("  %s %.1f", nm[i], h[i] / 256.0);
        printf("\n");
    }
    return 0;
}
