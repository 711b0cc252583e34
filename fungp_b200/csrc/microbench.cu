// Roofline denominators measured on the device the library runs on: register-resident DMMA (m8n8k4 f64) issue
// rate, plain DFMA issue rate, and a device-to-device copy.  MEASURED_PEAKS.json (driver-written) has no FP64
// entry, so the FP64 "tensor peak" the Cholesky fractions are quoted against is this DMMA loop.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

}  // namespace

int fgp_peaks_device(double* dmma_tflops, double* dfma_tflops, double* copy_gbs) {
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return FGP_ERR_CUDA_;
    const int nsm = prop.multiProcessorCount;
    double* dout = nullptr;
    if (cudaMalloc(&dout, 64) != cudaSuccess) return FGP_ERR_ALLOC_;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    // two CTAs of 8 warps per SM, as tools/dmma_sustained.cu: one warp per sub-partition already saturates the FP64
    // datapath (r01_dmma_sweep.log); four CTAs per SM only add power -- on a hot board the loop then came out at 29.7
    // instead of 37.1 TFLOP/s (power-capped clocks) while the real kernels, at 650-720 W, were unaffected
    const int grid = nsm * 2;
    {
        const int iters = 4096;
        dmma_peak_kernel<<<grid, 256>>>(dout, 64);   // warm-up
        double best = 0.0;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            dmma_peak_kernel<<<grid, 256>>>(dout, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            const double fl = (double)grid * 8.0 * iters * 16.0 * 512.0;   // warps * DMMAs * 2*8*8*4
            best = std::max(best, fl / (ms * 1e-3) / 1e12);
        }
        if (dmma_tflops) *dmma_tflops = best;
    }
    {
        const int iters = 8192;
        dfma_peak_kernel<<<grid, 256>>>(dout, 64);
        double best = 0.0;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            dfma_peak_kernel<<<grid, 256>>>(dout, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            const double fl = (double)grid * 256.0 * iters * 16.0 * 2.0;
            best = std::max(best, fl / (ms * 1e-3) / 1e12);
        }
        if (dfma_tflops) *dfma_tflops = best;
    }
    if (copy_gbs) {
        const size_t bytes = (size_t)1 << 30;
        char *a = nullptr, *b = nullptr;
        *copy_gbs = 0.0;
        if (cudaMalloc(&a, bytes) == cudaSuccess && cudaMalloc(&b, bytes) == cudaSuccess) {
            cudaMemset(a, 1, bytes);
            cudaMemcpy(b, a, bytes, cudaMemcpyDeviceToDevice);
            double best = 0.0;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, 0);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms, e0, e1);
                best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
            }
            *copy_gbs = best;
        }
        if (a) cudaFree(a);
        if (b) cudaFree(b);
        cudaGetLastError();
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dout);
    return cudaGetLastError() == cudaSuccess ? 0 : FGP_ERR_CUDA_;
}
