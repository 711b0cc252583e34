#include <cuda_runtime.h>
#include <cstdio>
template<int ILP>
__global__ void k(double* out, int iters) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
// variant: distinct A/B registers per DMMA (4 A x 8 B frags like the GEMM warp tile) + smem fragment loads
__global__ void kfrag(double* out, int iters, int useSmem) {
    extern __shared__ double sm[];
    double c[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const double* As = sm + ((threadIdx.x >> 5) & 3) * 1024 + (lane & 3) * 8 + (lane >> 2);
    const double* Bs = sm + 4096 + ((threadIdx.x >> 7) & 1) * 2048 + (lane & 3) * 8 + (lane >> 2);
    double a[4], b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = As[i * 256];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = Bs[j * 256];
    for (int it = 0; it < iters; ++it) {
        if (useSmem) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[i * 256 + kk * 32];
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = Bs[j * 256 + kk * 32];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 12345.678) out[0] = s;
}
template<int ILP> double run(int threads, int bps, int iters, double* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * bps;
    k<ILP><<<grid, threads>>>(d, 16);
    float best = 1e9;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<ILP><<<grid, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fl = (double)grid * (threads / 32) * (double)iters * ILP * 512.0;
    return fl / (best * 1e-3) / 1e12;
}
int main() {
    double* d; cudaMalloc(&d, 64);
    printf("ILP threads blocks/SM  TFLOPs\n");
    for (int threads : {32, 64, 128, 256, 512, 1024}) {
        printf("ilp4  %4d 1 %.2f\n", threads, run<4>(threads, 1, 8192, d));
        printf("ilp8  %4d 1 %.2f\n", threads, run<8>(threads, 1, 8192, d));
        printf("ilp16 %4d 1 %.2f\n", threads, run<16>(threads, 1, 4096, d));
        printf("ilp32 %4d 1 %.2f\n", threads, run<32>(threads, 1, 2048, d));
    }
    cudaFuncSetAttribute(kfrag, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int useSmem = 0; useSmem < 2; ++useSmem)
    for (int threads : {128, 256, 512}) {
        int iters = 512;
        kfrag<<<148, threads, 65536>>>(d, 4, useSmem);
        float best = 1e9;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0); kfrag<<<148, threads, 65536>>>(d, iters, useSmem); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double fl = 148.0 * (threads / 32) * (double)iters * 256 * 512.0;
        printf("frag smem=%d threads=%d: %.2f TF\n", useSmem, threads, fl / (best * 1e-3) / 1e12);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
