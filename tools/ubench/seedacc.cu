// Accuracy of the MUFU seeds rsqrt.approx.ftz.f64 / rcp.approx.ftz.f64 and of the Newton refinements used in sb2st.cu / sb2sb.cu.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(double* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // x sweeps [1, 4) (one octave pair for rsqrt) times assorted powers of two
    const double x = (1.0 + 3.0 * (i + 0.37) / n) * exp2((double)((i % 41) - 20) * 7.0);
    double ys, yr;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(ys) : "d"(x));
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(yr) : "d"(x));
    const double ts = 1.0 / sqrt(x), tr = 1.0 / x;
    double y1 = ys;
    { const double e = fma(-x * y1, y1, 1.0); y1 = fma(y1 * e, fma(e, 0.375, 0.5), y1); }
    double y2 = y1;
    { const double e = fma(-x * y2, y2, 1.0); y2 = fma(y2 * e, fma(e, 0.375, 0.5), y2); }
    double r1 = fma(yr, fma(-x, yr, 1.0), yr), r2 = fma(r1, fma(-x, r1, 1.0), r1), r3 = fma(r2, fma(-x, r2, 1.0), r2);
    out[8 * i + 0] = fabs(ys - ts) / ts; out[8 * i + 1] = fabs(y1 - ts) / ts; out[8 * i + 2] = fabs(y2 - ts) / ts;
    out[8 * i + 3] = fabs(yr - tr) / tr; out[8 * i + 4] = fabs(r1 - tr) / tr; out[8 * i + 5] = fabs(r2 - tr) / tr; out[8 * i + 6] = fabs(r3 - tr) / tr;
}
int main() {
    const int n = 1 << 22;
    double* d; cudaMalloc(&d, 8 * 8 * n);
    k<<<n / 256, 256>>>(d, n);
    double* h = new double[8 * n];
    cudaMemcpy(h, d, 8 * 8 * n, cudaMemcpyDeviceToHost);
    double m[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) for (int j = 0; j < 7; ++j) m[j] = fmax(m[j], h[8 * i + j]);
    printf("rsqrt seed %.3e (2^%.1f), 1 cubic step %.3e, 2 steps %.3e\n", m[0], log2(m[0]), m[1], m[2]);
    printf("rcp   seed %.3e (2^%.1f), 1 Newton %.3e, 2 %.3e, 3 %.3e\n", m[3], log2(m[3]), m[4], m[5], m[6]);
    return 0;
}
