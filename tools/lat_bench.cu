// dependent-chain latencies on one warp (clock64): DFMA, DMUL, DMMA, SHFL(double), LDS.64, rsqrt variants
#include <cuda_runtime.h>
#include <cstdio>
#define N 256
__device__ __forceinline__ double pivot_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}
__global__ void k(double* out, long long* cyc, double seed) {
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    sm[lane] = seed + lane; sm[lane + 32] = 1.0;
    __syncthreads();
    double x = seed, y = 1.0000001, z = 1e-9;
    long long t0, t1;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, y, z);
    t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = x * y;
    t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
    double c0 = x, c1 = x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(y), "d"(z));
    t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
    x += c0 + c1;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
    t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
    int idx = lane;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) { double v = sm[idx]; idx = ((int)v) & 31; }
    t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
    x += idx;
    x = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = rsqrt(x) + 2.0;
    t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = pivot_rsqrt(x) + 2.0;
    t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = sqrt(x) + 2.0;
    t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) x = 1.0 / x + 2.0;
    t1 = clock64(); if (lane == 0) cyc[8] = t1 - t0;
    // throughput of independent DFMA from one warp (16 chains)
    double a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = x + j;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < N / 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = fma(a[j], y, z);
    t1 = clock64(); if (lane == 0) cyc[9] = t1 - t0;
#pragma unroll
    for (int j = 0; j < 16; ++j) x += a[j];
    // broadcast LDS + FMA stream like potf2 (b): 32 independent accumulators, one LDS each
    double b[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = x + j;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 32; ++j) b[j] = fma(-y, sm[(i * 32 + j) & 63], b[j]);
    t1 = clock64(); if (lane == 0) cyc[10] = t1 - t0;
#pragma unroll
    for (int j = 0; j < 32; ++j) x += b[j];
    out[lane] = x;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 512); cudaMalloc(&c, 128);
    k<<<1, 32>>>(d, c, 1.5); k<<<1, 32>>>(d, c, 1.5); cudaDeviceSynchronize();
    long long h[11]; cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
    const char* nm[] = {"DFMA dep", "DMUL dep", "DMMA dep", "SHFL(f64) dep", "LDS.64 dep(+cvt)", "rsqrt()+add dep", "pivot_rsqrt+add dep", "sqrt()+add dep", "1/x+add dep", "DFMA 16 indep (per op)", "LDS bcast+DFMA 32 indep (per op)"};
    for (int i = 0; i < 11; ++i) printf("%-34s %.1f cycles\n", nm[i], (double)h[i] / (i == 9 ? (N / 4 * 16) : (i == 10 ? 256 : N)));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
