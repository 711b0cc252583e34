// Diagonal-block kernel of the blocked Cholesky: one launch factors the 128x128 block (L_kk) of every matrix of the
// batch and produces its inverse W_k = inv(L_kk).  The algorithm is in potf2_body.cuh.
#include "potf2_body.cuh"
#include <atomic>

namespace {
using namespace potf2;

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* M, long ld, long strideM, int k0, int nb, double* W, long strideW, int* info, int infoOff, int pdl) {
    extern __shared__ __align__(16) double sm[];
    const uint32_t barAddr = (uint32_t)__cvta_generic_to_shared(sm + POTF2_WORK_DOUBLES);
    init_barriers(barAddr);
    double* Mdiag = M + (long)blockIdx.x * strideM + (long)k0 + (long)k0 * ld;
    double* Wb = W + (long)blockIdx.x * strideW + (long)(k0 / NB) * NB * NB;
    // programmatic dependent launch: the barrier set-up above overlapped the previous kernel's tail
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    if (pdl == 2) asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    PROF(0);
    load_block(sm, Mdiag, ld, nb);
    factor_block(sm, barAddr, info ? info + blockIdx.x : nullptr, infoOff);
    store_block(sm, Mdiag, ld, nb, Wb);
}

}  // namespace

void launch_potf2(double* M, long ld, long strideM, int batch, int k0, int nb, double* W, long strideW, int* info,
                  int infoOff, cudaStream_t st, int pdl) {
    static std::atomic<bool> attr_set[FGP_MAXDEV];        // zero-initialised; several host threads may arrive at once
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < FGP_MAXDEV && !attr_set[dev]) {
        cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
        attr_set[dev] = true;
    }
    FGP_COUNT_LAUNCH();
    if (pdl) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(batch); cfg.blockDim = dim3(POTF2_THREADS); cfg.dynamicSmemBytes = POTF2_SMEM; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, potf2_inv_kernel, M, ld, strideM, k0, nb, W, strideW, info, infoOff, pdl);
        return;
    }
    potf2_inv_kernel<<<batch, POTF2_THREADS, POTF2_SMEM, st>>>(M, ld, strideM, k0, nb, W, strideW, info, infoOff, 0);
}
