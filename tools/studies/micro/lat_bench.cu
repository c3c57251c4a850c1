// Microbenchmarks of the instruction latencies / issue costs that shape the warp-per-controller kernel on B200 (sm_100a):
// dependent chains of DFMA, 64-bit shuffles, LDS, rsqrt / rcp, the FP64 MMA shapes (m8n8k4, m16n8k4, m16n8k8, m16n8k16),
// REDUX, and a shared-memory store -> __syncwarp -> load round trip.  One number per line: cycles per operation for a
// single warp (latency), and for 4 / 16 warps on one SM (aggregate issue cost = cycles * 1 / ops per warp).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_bench lat_bench.cu && ./lat_bench
#include <cstdio>
#include <cuda_runtime.h>

#define REP 512
__device__ __forceinline__ long long clk() { return clock64(); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], double a0, double a1, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void bench(int which, double seed, double* out, long long* cyc) {
    __shared__ double sm[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 3) & 1023);
    __syncthreads();
    double x = seed + lane * 1e-3, y = 1.0 + seed, acc = 0.0;
    double c[4] = {seed, seed, seed, seed}, d[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0}, f[4] = {0, 0, 0, 0};
    double a8[8], b4[4];
    for (int i = 0; i < 8; ++i) a8[i] = seed * (i + 1) * 1e-3;
    for (int i = 0; i < 4; ++i) b4[i] = seed * (i + 2) * 1e-3;
    int idx = lane;
    __syncthreads();
    const long long t0 = clk();
    switch (which) {
        case 0:
#pragma unroll 16
            for (int i = 0; i < REP; ++i) x = fma(x, y, seed);
            break;
        case 1:  // 4 independent DFMA chains
#pragma unroll 8
            for (int i = 0; i < REP; ++i) { c[0] = fma(c[0], y, seed); c[1] = fma(c[1], y, seed); c[2] = fma(c[2], y, seed); c[3] = fma(c[3], y, seed); }
            x = c[0] + c[1] + c[2] + c[3];
            break;
        case 2:
#pragma unroll 16
            for (int i = 0; i < REP; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
            break;
        case 3: {
            float xf = (float)x;
#pragma unroll 16
            for (int i = 0; i < REP; ++i) xf = __shfl_xor_sync(0xffffffffu, xf, 1) + 1.0f;
            x = xf;
        } break;
        case 4:  // LDS.64 pointer chase
#pragma unroll 16
            for (int i = 0; i < REP; ++i) idx = (int)sm[idx];
            x = idx;
            break;
        case 5:
#pragma unroll 8
            for (int i = 0; i < REP; ++i) x = rsqrt(x) + 1.0;
            break;
        case 6:
#pragma unroll 8
            for (int i = 0; i < REP; ++i) x = __drcp_rn(x) + 1.0;
            break;
        case 7:
#pragma unroll 8
            for (int i = 0; i < REP; ++i) x = 1.0 / x + 1.0;
            break;
        case 8:
#pragma unroll 8
            for (int i = 0; i < REP; ++i) x = sqrt(x) + 1.0;
            break;
        case 9:  // DMMA m8n8k4 dependent chain
#pragma unroll 8
            for (int i = 0; i < REP; ++i) dmma884(c[0], c[1], x, y);
            x = c[0] + c[1];
            break;
        case 10:  // 3 independent m8n8k4 chains (as in the Phi build)
#pragma unroll 4
            for (int i = 0; i < REP; ++i) { dmma884(c[0], c[1], x, y); dmma884(d[0], d[1], x, y); dmma884(e[0], e[1], y, x); }
            x = c[0] + c[1] + d[0] + d[1] + e[0] + e[1];
            break;
        case 11:
#pragma unroll 8
            for (int i = 0; i < REP; ++i) dmma1684(c, x, y, y);
            x = c[0] + c[1] + c[2] + c[3];
            break;
        case 12: {
            double a4[4] = {x, y, x + 1, y + 1}, b2[2] = {y, x};
#pragma unroll 8
            for (int i = 0; i < REP; ++i) dmma1688(c, a4, b2);
            x = c[0] + c[1] + c[2] + c[3];
        } break;
        case 13:
#pragma unroll 8
            for (int i = 0; i < REP; ++i) dmma16816(c, a8, b4);
            x = c[0] + c[1] + c[2] + c[3];
            break;
        case 14:  // 2 independent m16n8k16 chains
#pragma unroll 4
            for (int i = 0; i < REP; ++i) { dmma16816(c, a8, b4); dmma16816(d, a8, b4); }
            x = c[0] + c[1] + c[2] + c[3] + d[0] + d[1] + d[2] + d[3];
            break;
        case 15: {  // REDUX max u32 chain
            unsigned u = (unsigned)lane + (unsigned)seed;
#pragma unroll 16
            for (int i = 0; i < REP; ++i) u = __reduce_max_sync(0xffffffffu, u) + (unsigned)lane;
            x = u;
        } break;
        case 16:  // STS -> syncwarp -> LDS (other lane's slot)
#pragma unroll 8
            for (int i = 0; i < REP; ++i) {
                sm[(threadIdx.x & ~31) + lane] = x;
                __syncwarp();
                x = sm[(threadIdx.x & ~31) + (lane ^ 1)] + 1.0;
                __syncwarp();
            }
            break;
        case 17:  // 64-bit butterfly sum (5 steps) chain
#pragma unroll 4
            for (int i = 0; i < REP; ++i) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                x *= 1e-2;
            }
            break;
        case 18:  // dependent DFMA + 64-bit broadcast (one substitution step)
#pragma unroll 16
            for (int i = 0; i < REP; ++i) x = fma(-y, __shfl_sync(0xffffffffu, x, i & 15), x);
            break;
        case 19:  // 2 independent m16n8k8 chains
        {
            double a4[4] = {x, y, x + 1, y + 1}, b2[2] = {y, x};
#pragma unroll 4
            for (int i = 0; i < REP; ++i) { dmma1688(c, a4, b2); dmma1688(d, a4, b2); }
            x = c[0] + c[1] + c[2] + c[3] + d[0] + d[1] + d[2] + d[3];
        } break;
        case 20:  // LDS.128 broadcast loads feeding DFMA (row products)
#pragma unroll 8
            for (int i = 0; i < REP; ++i) {
                const double2 v = *reinterpret_cast<const double2*>(&sm[(i * 2) & 1022]);
                acc = fma(v.x, x, acc);
                acc = fma(v.y, y, acc);
            }
            x = acc;
            break;
    }
    const long long t1 = clk();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + f[0];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const char* names[] = {"DFMA dependent chain", "DFMA x4 independent (per group of 4)", "SHFL 64-bit + DADD chain", "SHFL 32-bit + FADD chain",
                           "LDS.64 pointer chase (+cvt)", "rsqrt(double) + DADD chain", "__drcp_rn + DADD chain", "1.0/x + DADD chain",
                           "sqrt(double) + DADD chain", "DMMA m8n8k4 dependent", "DMMA m8n8k4 x3 independent (per group of 3)",
                           "DMMA m16n8k4 dependent", "DMMA m16n8k8 dependent", "DMMA m16n8k16 dependent", "DMMA m16n8k16 x2 independent (per pair)",
                           "REDUX.MAX.U32 + IADD chain", "STS -> syncwarp -> LDS -> syncwarp round trip", "64-bit butterfly sum (5 steps) + DMUL",
                           "substitution step: SHFL.IDX 64-bit + DFMA", "DMMA m16n8k8 x2 independent (per pair)", "LDS.128 + 2 DFMA (per pair)"};
    for (int which = 0; which <= 20; ++which) {
        printf("%-50s", names[which]);
        for (int threads : {32, 128, 512}) {
            bench<<<1, threads>>>(which, 1.25, out, cyc);
            cudaDeviceSynchronize();
            bench<<<1, threads>>>(which, 1.25, out, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            long long c = 0;
            cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) printf("  ERR(%s)", cudaGetErrorString(e));
            printf("  %3d thr: %7.1f cyc/op", threads, (double)c / REP);
        }
        printf("\n");
    }
    return 0;
}
