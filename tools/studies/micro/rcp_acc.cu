// Accuracy of the hardware reciprocal seed (rcp.approx.ftz.f64 = MUFU.RCP64H) and of rcp_fast's single third-order step
// (bmpc_warp.cuh), over 2^24 positive doubles spread over 600 binades:  nvcc -arch=sm_100a -o rcp_acc rcp_acc.cu && ./rcp_acc
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ double seed(double d) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r; }
__device__ __forceinline__ double rcp_fast(double d) { double r = seed(d); const double e = fma(-d, r, 1.0); return fma(r, fma(e, e, e), r); }
__global__ void k(double* out) {
    double m0 = 0, m1 = 0, m2 = 0;
    unsigned long long x = 0x9E3779B97F4A7C15ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    for (int i = 0; i < 4096; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const double mant = 1.0 + (double)(x >> 11) * (1.0 / 9007199254740992.0);
        const double d = ldexp(mant, (int)(x % 600) - 300);
        const double ex = 1.0 / d;  // correctly rounded
        m0 = fmax(m0, fabs(seed(d) - ex) / ex);
        m1 = fmax(m1, fabs(rcp_fast(d) - ex) / ex);
        m2 = fmax(m2, fabs(fma(-d, rcp_fast(d), 1.0)));
    }
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(m0));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(m1));
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(m2));
}
int main() {
    double* d; cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
    k<<<64, 64>>>(d);
    double h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("max relative error: seed %.3e (2^%.1f)   rcp_fast %.3e (%.2f ulp)   |1 - d*rcp_fast(d)| %.3e\n", h[0], log2(h[0]), h[1], h[1] / 1.1102230246251565e-16, h[2]);
    return cudaGetLastError() != cudaSuccess;
}
