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

// 1/sqrt(d) for the pivots: MUFU seed (rsqrt.approx, ~2^-20 relative) + one cubically convergent correction, all
// FMA.  This sits on the factorisation's only truly serial chain (one pivot per column), so its latency -- not its
// throughput -- is what matters; the libdevice rsqrt() is ~2x longer.  Result within ~1 ulp.
__device__ __forceinline__ double pivot_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d * y, y, 1.0);                 // 1 - d y^2
    return fma(y * e, fma(0.375, e, 0.5), y);             // y (1 + e/2 + 3 e^2/8)
}

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
#pragma unroll 16
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
                // lane = row.  Two dependency chains per column, both kept in registers / shuffles:
                //   pivots:  dg (this lane's own diagonal entry, updated with -l^2 every column) -> shfl -> rsqrt
                //   column:  a[j+1] gets its last update from a shuffle of l, not from shared memory
                // the remaining updates a[c], c > j+1, read column j back from shared memory (broadcast) off-chain.
                double a[SB];
#pragma unroll
                for (int j = 0; j < SB; ++j) a[j] = T[(o + j) * LDT + o + lane];
                double dg = T[(o + lane) * LDT + o + lane];
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    double d = __shfl_sync(0xffffffffu, dg, j);
                    if (!(d > 0.0)) {   // not positive definite: remember the minor's order, keep going finite
                        if (lane == 0 && info && *(info + blockIdx.x) == 0) *(info + blockIdx.x) = infoOff + o + j + 1;
                        d = 1.0;
                    }
                    const double rinv = pivot_rsqrt(d);
                    double l = 0.0;
                    if (lane > j) l = a[j] * rinv;
                    else if (lane == j) { l = d * rinv; rinvL[j] = rinv; }
                    dg = fma(-l, l, dg);
                    if (lane >= j) T[(o + j) * LDT + o + lane] = l;
                    if (j + 1 < SB) {
                        const double ln = __shfl_sync(0xffffffffu, l, j + 1);
                        a[j + 1] = fma(-l, ln, a[j + 1]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int c = j + 2; c < SB; ++c) a[c] = fma(-l, T[(o + j) * LDT + o + c], a[c]);
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
        const int ncg = (NB - c_first) / 4;   // column groups of 4; one warp per group, lanes own rows lane + 32 i
        for (int q = warp; q < ncg; q += POTF2_THREADS / 32) {
            const int c0 = c_first + q * 4;
            double acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 8
            for (int k = 0; k < SB; ++k) {
                double xr[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) xr[i] = XP[k * NB + lane + 32 * i];          // conflict-free
                const double2 xc01 = *reinterpret_cast<const double2*>(&XP[k * NB + c0]);      // broadcast
                const double2 xc23 = *reinterpret_cast<const double2*>(&XP[k * NB + c0 + 2]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i][0] = fma(xr[i], xc01.x, acc[i][0]);
                    acc[i][1] = fma(xr[i], xc01.y, acc[i][1]);
                    acc[i][2] = fma(xr[i], xc23.x, acc[i][2]);
                    acc[i][3] = fma(xr[i], xc23.y, acc[i][3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = lane + 32 * i, c = c0 + j;
                    if (r >= c || r < c_first) T[c * LDT + r] -= acc[i][j];
                }
        }
        __syncthreads();
    }

    // ---- write back L (lower triangle of the valid part) and W = inv(L) (128x128, unit padded)
#pragma unroll 8
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
    FGP_COUNT_LAUNCH();
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
    FGP_COUNT_LAUNCH();
    potf2_inv_kernel<false><<<ntiles, POTF2_THREADS, POTF2_SMEM, st>>>(const_cast<double*>(M), ld, 0, 0, 0, ncolsValid, W,
                                                                       0, nullptr, 0);
}
