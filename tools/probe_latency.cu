// Dependent-issue latencies of the FP64 / shuffle / shared-memory operations the latency-bound kernels (sb2st, KPM Lanczos,
// panel QR) are made of: one warp, one chain, clock64 around 1024 dependent operations.  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void lat(double* out, long long* cyc, double a, double b) {
    __shared__ double sm[64];
    sm[threadIdx.x & 63] = a;
    __syncthreads();
    double x = a + threadIdx.x * 1e-9;
    int idx = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) {
        if (OP == 0) x = fma(x, b, a);
        if (OP == 1) x = x + b;
        if (OP == 2) x = x * b;
        if (OP == 3) x += __shfl_xor_sync(0xffffffffu, x, 1);
        if (OP == 4) { idx = (int)sm[idx] + (threadIdx.x & 31); }
        if (OP == 5) x = sqrt(x + 2.0);
        if (OP == 6) x = 1.0 / (x + 2.0);
        if (OP == 7) x = rsqrt(x + 2.0);
        if (OP == 8) { x = fma(x, b, a); if (x == 0.0) x = -1e-300; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
    out[threadIdx.x] = x + idx;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 64); cudaMalloc(&cyc, 8 * 16);
    const char* names[] = {"DFMA", "DADD", "DMUL", "SHFL.64 + DADD", "LDS.64 -> address", "sqrt(double)", "1/x (double)", "rsqrt(double)", "DFMA + zero test/select"};
    lat<0><<<1, 32>>>(out, cyc, 0.0, 0.999); lat<1><<<1, 32>>>(out, cyc, 0.0, 1e-3); lat<2><<<1, 32>>>(out, cyc, 1.0, 0.999);
    lat<3><<<1, 32>>>(out, cyc, 1.0, 0.999); lat<4><<<1, 32>>>(out, cyc, 0.0, 0.999); lat<5><<<1, 32>>>(out, cyc, 1.0, 0.999);
    lat<6><<<1, 32>>>(out, cyc, 1.0, 0.999); lat<7><<<1, 32>>>(out, cyc, 1.0, 0.999); lat<8><<<1, 32>>>(out, cyc, 0.5, 0.999);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 8 * 16, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 9; ++i) printf("%-26s %6.1f cycles per dependent op\n", names[i], h[i] / 1024.0);
    return 0;
}
