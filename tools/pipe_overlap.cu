// Do DMMA (mma.sync m8n8k4 f64) and DFMA share one issue pipe on sm_100a?  Three runs of the same kernel:
// DMMA only, DFMA only, both interleaved in one instruction stream.  If the mixed time is close to the sum the
// two share the FP64 datapath; if it is close to the max they are separate pipes.
#include <cuda_runtime.h>
#include <cstdio>
template <int NM, int NF>
__global__ void __launch_bounds__(256) k(double* out, int iters) {
    double c[8][2], f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < NM)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (2 * i + j < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(f[2 * i + j]) : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    if (s == 12345.678) out[0] = s;
}
template <int NM, int NF>
float run(double* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NM, NF><<<148 * 2, 256>>>(d, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<NM, NF><<<148 * 2, 256>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warps = 148.0 * 2 * 8;
    printf("DMMA/iter=%d DFMA/iter=%2d : %8.3f ms   DMMA %.2f TF/s   DFMA %.2f TF/s\n", NM, NF, ms,
           warps * iters * NM * 512.0 / (ms * 1e-3) / 1e12, warps * iters * NF * 64.0 / (ms * 1e-3) / 1e12);
    return ms;
}
int main() {
    double* d; cudaMalloc(&d, 64);
    const int iters = 40000;
    float m = run<8, 0>(d, iters);
    float f = run<0, 16>(d, iters);
    float b = run<8, 16>(d, iters);
    float b2 = run<8, 8>(d, iters);
    float f8 = run<0, 8>(d, iters);
    printf("mixed(8,16)/(dmma+dfma16) = %.3f   mixed/max = %.3f\n", b / (m + f), b / (m > f ? m : f));
    printf("mixed(8,8)/(dmma+dfma8)  = %.3f   mixed/max = %.3f\n", b2 / (m + f8), b2 / (m > f8 ? m : f8));
    return 0;
}
