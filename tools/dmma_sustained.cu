// sustained DMMA rate: back-to-back launches for ~4 s, rate per 250 ms window
#include <cuda_runtime.h>
#include <cstdio>
#include <chrono>
__global__ void k(double* out, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
int main() {
    double* d; cudaMalloc(&d, 64);
    const int grid = 148 * 2, threads = 256, iters = 20000;
    const double fl = (double)grid * (threads / 32) * (double)iters * 16 * 512.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<grid, threads>>>(d, 16); cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int w = 0; w < 16; ++w) {
        cudaEventRecord(e0);
        int nl = 0;
        auto ws = std::chrono::steady_clock::now();
        while (std::chrono::duration<double>(std::chrono::steady_clock::now() - ws).count() < 0.25) { k<<<grid, threads>>>(d, iters); ++nl; cudaStreamSynchronize(0); }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("t=%.2fs  %.2f TF/s (%d launches)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), fl * nl / (ms * 1e-3) / 1e12, nl);
        fflush(stdout);
    }
    return 0;
}
