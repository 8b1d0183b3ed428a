// Are the registers of warps that exit early returned to the SM before their CTA ends?  12 warps x 96 registers = 36.9k registers per
// CTA: two CTAs only fit an SM (65.5k) if the three warps that exit at once give their registers back.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) k(int* maxc, int* cur, double* out, int exit_early) {
    const int w = threadIdx.x >> 5;
    if (exit_early && (w == 0 || w == 4 || w == 8)) return;
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    // force ~96 registers live
    double r[40];
    for (int i = 0; i < 40; ++i) r[i] = threadIdx.x * 0.5 + i;
    if (threadIdx.x == 32) { int c = atomicAdd(&cur[smid], 1) + 1; atomicMax(&maxc[smid], c); }
    long long t0 = clock64();
    while (clock64() - t0 < 2000000) {
#pragma unroll
        for (int i = 0; i < 40; ++i) r[i] = fma(r[i], 1.0000001, r[(i + 7) % 40]);
    }
    if (threadIdx.x == 32) atomicAdd(&cur[smid], -1);
    double s = 0; for (int i = 0; i < 40; ++i) s += r[i];
    out[blockIdx.x * 384 + threadIdx.x] = s;
}
int main() {
    int *maxc, *cur; double* out;
    cudaMalloc(&maxc, 4 * 256); cudaMalloc(&cur, 4 * 256); cudaMalloc(&out, 8 * 384 * 1024);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k); printf("registers per thread: %d\n", fa.numRegs);
    for (int e = 0; e < 2; ++e) {
        cudaMemset(maxc, 0, 1024); cudaMemset(cur, 0, 1024);
        k<<<592, 384>>>(maxc, cur, out, e); cudaDeviceSynchronize();
        int h[256]; cudaMemcpy(h, maxc, 1024, cudaMemcpyDeviceToHost);
        int m = 0; for (int i = 0; i < 256; ++i) m = h[i] > m ? h[i] : m;
        printf("exit_early=%d: max CTAs concurrently on one SM = %d\n", e, m);
    }
    return 0;
}
