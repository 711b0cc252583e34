// Is one FP64 DMMA (mma.sync m8n8k4) the same as four chained FMAs in k order?  Random operands, bitwise comparison.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
__global__ void k(const double* A, const double* B, const double* C, double* D, double* E, int mode) {
    // A: 8x4 row-major (row = lane>>2, k = lane&3); B: 4x8 (k = lane&3, n = lane>>2); C/D: 8x8, lane holds (row = lane>>2, cols 2*(lane&3), +1)
    const int lane = threadIdx.x;
    double a = A[blockIdx.x * 32 + (lane >> 2) * 4 + (lane & 3)];
    double b = B[blockIdx.x * 32 + (lane & 3) * 8 + (lane >> 2)];
    double c0 = C[blockIdx.x * 64 + (lane >> 2) * 8 + 2 * (lane & 3)], c1 = C[blockIdx.x * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + 1];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    D[blockIdx.x * 64 + (lane >> 2) * 8 + 2 * (lane & 3)] = c0;
    D[blockIdx.x * 64 + (lane >> 2) * 8 + 2 * (lane & 3) + 1] = c1;
    // reference chains
    for (int h = 0; h < 2; ++h) {
        const int r = lane >> 2, n = 2 * (lane & 3) + h;
        double s = C[blockIdx.x * 64 + r * 8 + n];
        if (mode == 0) for (int kk = 0; kk < 4; ++kk) s = fma(A[blockIdx.x * 32 + r * 4 + kk], B[blockIdx.x * 32 + kk * 8 + n], s);
        else for (int kk = 3; kk >= 0; --kk) s = fma(A[blockIdx.x * 32 + r * 4 + kk], B[blockIdx.x * 32 + kk * 8 + n], s);
        E[blockIdx.x * 64 + r * 8 + n] = s;
    }
}
int main() {
    const int NB = 4096;
    double *hA = (double*)malloc(NB * 32 * 8), *hB = (double*)malloc(NB * 32 * 8), *hC = (double*)malloc(NB * 64 * 8), *hD = (double*)malloc(NB * 64 * 8), *hE = (double*)malloc(NB * 64 * 8);
    srand(5);
    auto rnd = [](int spread) { return ((double)rand() / RAND_MAX - 0.5) * pow(2.0, rand() % spread - spread / 2); };
    for (int i = 0; i < NB * 32; ++i) { hA[i] = rnd(20); hB[i] = rnd(20); }
    for (int i = 0; i < NB * 64; ++i) hC[i] = rnd(30);
    double *A, *B, *C, *D, *E;
    cudaMalloc(&A, NB * 32 * 8); cudaMalloc(&B, NB * 32 * 8); cudaMalloc(&C, NB * 64 * 8); cudaMalloc(&D, NB * 64 * 8); cudaMalloc(&E, NB * 64 * 8);
    cudaMemcpy(A, hA, NB * 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, NB * 32 * 8, cudaMemcpyHostToDevice); cudaMemcpy(C, hC, NB * 64 * 8, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
        k<<<NB, 32>>>(A, B, C, D, E, mode);
        cudaMemcpy(hD, D, NB * 64 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hE, E, NB * 64 * 8, cudaMemcpyDeviceToHost);
        long diff = 0;
        for (int i = 0; i < NB * 64; ++i) diff += memcmp(&hD[i], &hE[i], 8) != 0;
        printf("DMMA vs FMA chain k %s: %ld of %d entries differ (%s)\n", mode == 0 ? "0..3" : "3..0", diff, NB * 64, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
