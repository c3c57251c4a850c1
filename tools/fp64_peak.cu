// Measures the B200 FP64 (DFMA) vector-pipe peak and a shared-memory-free ILP sweep.
// Used once per pool to set the roofline denominator for the fp64-bound solver kernels
// (MEASURED_PEAKS.json has no fp64 entry). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
double run(int blocks_per_sm, int nsm, int iters) {
    double* out;
    cudaMalloc(&out, sizeof(double) * 1024 * 1024);
    int grid = blocks_per_sm * nsm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) dfma_kernel<ILP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 10; ++r) {
        cudaEventRecord(e0);
        dfma_kernel<ILP><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double flops = 2.0 * ILP * (double)iters * 256.0 * grid;
    cudaFree(out);
    return flops / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, nsm, p.clockRate);
    printf(", \"dfma_tflops\": {");
    printf("\"ilp1_occ8\": %.2f", run<1>(8, nsm, 20000));
    printf(", \"ilp4_occ8\": %.2f", run<4>(8, nsm, 8000));
    printf(", \"ilp8_occ4\": %.2f", run<8>(4, nsm, 8000));
    printf(", \"ilp8_occ8\": %.2f", run<8>(8, nsm, 8000));
    printf(", \"ilp8_occ1\": %.2f", run<8>(1, nsm, 8000));
    printf("}}\n");
    return 0;
}
