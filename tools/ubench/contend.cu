// FP64 pipe arbitration: a victim warp runs a dependent DFMA chain while aggressor warps stream independent DMMAs.
// Which warps share a sub-partition (warp index mod 4?) and how long each victim instruction waits.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* t, int victim, unsigned aggr_mask, int* wid) {
    const int w = threadIdx.x >> 5;
    unsigned hwid; asm volatile("mov.u32 %0, %%warpid;" : "=r"(hwid));
    if ((threadIdx.x & 31) == 0) wid[w] = hwid;
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    double x = 1.0 + threadIdx.x, y = 1.0000001;
    if (w == victim) {
        long long a = clock64();
#pragma unroll 16
        for (int i = 0; i < 4096; ++i) x = fma(x, y, 0.5);
        long long b = clock64();
        if ((threadIdx.x & 31) == 0) t[0] = b - a;
        stop = 1;
    } else if (aggr_mask >> w & 1) {
        double d[16];
        for (int i = 0; i < 16; ++i) d[i] = x + i;
        while (!stop) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma(d[2 * j], d[2 * j + 1], y, y);
        }
        for (int i = 0; i < 16; ++i) x += d[i];
    }
    out[threadIdx.x] = x;
}
int main() {
    double* o; long long* t; int* wid; cudaMalloc(&o, 8 * 1024); cudaMalloc(&t, 64); cudaMalloc(&wid, 64 * 4);
    struct { int victim; unsigned mask; const char* what; } cases[] = {
        {8, 0x000, "no aggressors"}, {8, 0x001, "warp 0 (same mod 4)"}, {8, 0x011, "warps 0,4 (same mod 4)"}, {8, 0x002, "warp 1"},
        {8, 0x0ee, "warps 1,2,3,5,6,7"}, {8, 0x0ff, "warps 0..7"}, {8, 0xeee, "warps 1,2,3,5,6,7,9,10,11"}};
    for (auto& c : cases) {
        k<<<1, 32 * 12>>>(o, t, c.victim, c.mask, wid); cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
        int hw[12]; cudaMemcpy(hw, wid, 48, cudaMemcpyDeviceToHost);
        printf("victim warp %d (hw %d), aggressors %-28s: %.1f cycles per dependent DFMA\n", c.victim, hw[c.victim], c.what, h / 4096.0);
    }
    return 0;
}
