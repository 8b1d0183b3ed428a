// DMMA.8x8x4 throughput against resident warps per SM sub-partition and independent accumulators per warp
// (developer probe: what the FP64 tensor pipe can take from 1, 2, 3, 4 ... warps).  nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC> __global__ void tput(double* out, int iters) {
    double seed = threadIdx.x * 1e-3;
    double c[NACC][2];
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0;
    double a = seed, b = 1.0 - seed;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) mma884(c[i][0], c[i][1], a, b);
    double s = 0;
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
template <int NACC> void run(int warps_per_sm, int nsm) {
    double* out; cudaMalloc(&out, 8);
    int iters = 20000;
    tput<NACC><<<nsm, warps_per_sm * 32>>>(out, 100); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); tput<NACC><<<nsm, warps_per_sm * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = double(nsm) * warps_per_sm * iters * NACC * 512.0;
    printf("warps/SM=%2d (%.2g per sub-partition) acc=%2d : %.2f TFLOP/s\n", warps_per_sm, warps_per_sm / 4.0, NACC, flops / best * 1e-9);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    for (int w : {4, 8, 12, 16, 32}) { run<1>(w, nsm); run<2>(w, nsm); run<4>(w, nsm); run<8>(w, nsm); run<16>(w, nsm); }
    return 0;
}
