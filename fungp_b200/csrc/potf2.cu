// Diagonal-block kernel of the blocked Cholesky: factor one 128x128 block (L_kk) AND produce its inverse
// W_k = inv(L_kk), which turns the panel triangular solve into a DMMA GEMM (gemm_dmma.cu).
// Replaces the unblocked part of base::chol / dpotrf (R/3_training_SF.R:249, R/4_prediction_SF.R:28,
// R/5_simulation_SF.R:14) and of backsolve / dtrsm (R/3_training_SF.R:267, R/4_prediction_SF.R:9,29).
//
// One CTA per matrix.  The block lives in shared memory as a 128 x 129 column-major array T:
//   lower triangle (r >= c): the block being factored -> L
//   strict upper   (r <  c): E(r,c) = (L^-T)(r,c) = W(c,r); E starts as the identity (diagonal kept in dinvE)
// Four sub-steps of 32 columns: (a) warp 0 factors the 32x32 diagonal sub-block, one lane per row, right-looking,
// column broadcast through shared memory; (b) 128 threads, one per row, forward-substitute their 32 entries
// (rows below the sub-block and the E rows); (c) all threads apply the rank-32 update to the columns to the right.
#include "common.cuh"

namespace {

constexpr int NB = 128;
constexpr int LDT = 129;
constexpr int SB = 32;
constexpr int POTF2_THREADS = 256;
constexpr int POTF2_SMEM = (NB * LDT + NB * SB + SB + NB) * 8;

template <bool FACTOR>
__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* M, long ld, long strideM, int k0_first, int nb_first, int ncolsValid, double* W,
                 long strideW, int* info, int infoOff) {
    extern __shared__ __align__(16) double sm[];
    double* T = sm;                       // 128 x 129
    double* XP = sm + NB * LDT;           // 128 x 32 staged panel (col-major, ld 128)
    double* rinvL = XP + NB * SB;         // 32: 1/L_jj of the current sub-block
    double* dinvE = rinvL + SB;           // 128: diagonal of L^-1
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    // FACTOR: blockIdx.x = batch element, one tile (k0_first, nb_first)
    // !FACTOR: blockIdx.x = tile index (all already factored), single matrix
    int k0, nb;
    double* Mb;
    double* Wb;
    if (FACTOR) {
        k0 = k0_first;
        nb = nb_first;
        Mb = M + (long)blockIdx.x * strideM;
        Wb = W + (long)blockIdx.x * strideW + (long)(k0 / NB) * NB * NB;
    } else {
        k0 = blockIdx.x * NB;
        nb = min(NB, ncolsValid - k0);
        Mb = M;
        Wb = W + (long)blockIdx.x * NB * NB;
    }

    // ---- load: lower triangle of the block, identity padding, zeros above the diagonal
    for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
        const int r = idx & (NB - 1), c = idx >> 7;
        double v = 0.0;
        if (r >= c) {
            if (r < nb && c < nb) v = Mb[(long)(k0 + r) + (long)(k0 + c) * ld];
            else if (r == c) v = 1.0;
        }
        T[c * LDT + r] = v;
    }
    __syncthreads();

    for (int s = 0; s < NB / SB; ++s) {
        const int o = s * SB;
        // ---- (a) factor the 32x32 diagonal sub-block (warp 0), or just invert its diagonal
        if (warp == 0) {
            if (FACTOR) {
                double a[SB];
#pragma unroll
                for (int j = 0; j < SB; ++j) a[j] = T[(o + j) * LDT + o + lane];
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    double d = __shfl_sync(0xffffffffu, a[j], j);
                    if (!(d > 0.0)) {   // not positive definite: remember the minor's order, keep going finite
                        if (lane == 0 && info && *(info + blockIdx.x) == 0) *(info + blockIdx.x) = infoOff + o + j + 1;
                        d = 1.0;
                    }
                    const double rinv = rsqrt(d);
                    double l = 0.0;
                    if (lane > j) l = a[j] * rinv;
                    else if (lane == j) { l = d * rinv; rinvL[j] = rinv; }
                    if (lane >= j) T[(o + j) * LDT + o + lane] = l;
                    __syncwarp();
#pragma unroll
                    for (int c = j + 1; c < SB; ++c) a[c] = fma(-l, T[(o + j) * LDT + o + c], a[c]);
                }
            } else {
                rinvL[lane] = 1.0 / T[(o + lane) * LDT + o + lane];
            }
        }
        __syncthreads();
        // ---- (b) forward substitution, one thread per row: x L_ss' = b
        if (tid < NB) {
            const int r = tid;
            const bool diagrow = (r >= o) && (r < o + SB);
            double bvec[SB];
#pragma unroll
            for (int j = 0; j < SB; ++j) bvec[j] = diagrow ? ((r - o == j) ? 1.0 : 0.0) : T[(o + j) * LDT + r];
#pragma unroll
            for (int k = 0; k < SB; ++k) {
                const double x = bvec[k] * rinvL[k];
                bvec[k] = x;
#pragma unroll
                for (int c = k + 1; c < SB; ++c) bvec[c] = fma(-x, T[(o + k) * LDT + o + c], bvec[c]);
            }
#pragma unroll
            for (int k = 0; k < SB; ++k) {
                XP[k * NB + r] = bvec[k];
                if (!diagrow) T[(o + k) * LDT + r] = bvec[k];
                else if (k > r - o) T[(o + k) * LDT + r] = bvec[k];
                else if (k == r - o) dinvE[r] = bvec[k];
            }
        }
        __syncthreads();
        // ---- (c) rank-32 update of the columns to the right: T(r,c) -= sum_k XP(r,k) XP(c,k)
        //      for (r >= c) [block being factored] or (r < o+32) [E rows already started]
        const int c_first = o + SB;
        const int ncg = (NB - c_first) / 4;   // column groups of 4
        for (int q = tid; q < 32 * ncg; q += POTF2_THREADS) {
            const int r0 = (q & 31) * 4;
            const int c0 = c_first + (q >> 5) * 4;
            if (r0 + 3 < c0 && r0 >= c_first) continue;   // wholly inside the not-yet-started E region
            double acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 8
            for (int k = 0; k < SB; ++k) {
                double xr[4], xc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) xr[i] = XP[k * NB + r0 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) xc[j] = XP[k * NB + c0 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fma(xr[i], xc[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = r0 + i, c = c0 + j;
                    if (r >= c || r < c_first) T[c * LDT + r] -= acc[i][j];
                }
        }
        __syncthreads();
    }

    // ---- write back L (lower triangle of the valid part) and W = inv(L) (128x128, unit padded)
    for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
        const int r = idx & (NB - 1), c = idx >> 7;
        if (FACTOR && r >= c && r < nb && c < nb) Mb[(long)(k0 + r) + (long)(k0 + c) * ld] = T[c * LDT + r];
        double w = 0.0;
        if (r > c) w = T[r * LDT + c];          // W(r,c) = E(c,r), stored at T(row c, col r)
        else if (r == c) w = dinvE[r];
        Wb[idx] = w;
    }
}

}  // namespace

void launch_potf2(double* M, long ld, long strideM, int batch, int k0, int nb, double* W, long strideW, int* info,
                  int infoOff, cudaStream_t st) {
    static bool attr_set[FGP_MAXDEV] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < FGP_MAXDEV && !attr_set[dev]) {
        cudaFuncSetAttribute(potf2_inv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
        cudaFuncSetAttribute(potf2_inv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
        attr_set[dev] = true;
    }
    potf2_inv_kernel<true><<<batch, POTF2_THREADS, POTF2_SMEM, st>>>(M, ld, strideM, k0, nb, 0, W, strideW, info, infoOff);
}

void launch_trtri_tiles(const double* M, long ld, int ntiles, int ncolsValid, double* W, cudaStream_t st) {
    static bool attr_set[FGP_MAXDEV] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < FGP_MAXDEV && !attr_set[dev]) {
        cudaFuncSetAttribute(potf2_inv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
        cudaFuncSetAttribute(potf2_inv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
        attr_set[dev] = true;
    }
    potf2_inv_kernel<false><<<ntiles, POTF2_THREADS, POTF2_SMEM, st>>>(const_cast<double*>(M), ld, 0, 0, 0, ncolsValid, W,
                                                                       0, nullptr, 0);
}
