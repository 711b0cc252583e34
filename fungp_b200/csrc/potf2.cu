// Diagonal-block kernel of the blocked Cholesky: factor one 128x128 block (L_kk) AND produce its inverse
// W_k = inv(L_kk), which turns the panel triangular solve into a DMMA GEMM (gemm_dmma.cu).
// Replaces the unblocked part of base::chol / dpotrf (R/3_training_SF.R:249, R/4_prediction_SF.R:28,
// R/5_simulation_SF.R:14) and of backsolve / dtrsm (R/3_training_SF.R:267, R/4_prediction_SF.R:9,29).
//
// One CTA (8 warps) per matrix; this is the serial chain of the factorisation, so the design goal is latency:
// shared-memory loads cost ~7 cycles of issue each on this part (measured, tools/lat_bench.cu), FP64 FMAs have 8
// cycles of latency and a DMMA 26, hence: everything that is a matrix product goes through DMMA fragments, and the
// only scalar code is the 32-column chain of step (a).
//
// The block lives in shared memory as a 128 x 132 column-major array T:
//   lower triangle (r >= c): the block being factored -> L
//   strict upper   (r <  c): E(r,c) = (L^-T)(r,c) = W(c,r); E starts as the identity (its diagonal is kept in dinvE)
// Four sub-steps s of 32 columns (o = 32 s):
//   (a) warp 0, lane = row: right-looking Cholesky of the 32x32 diagonal sub-block AND, carried along with the same
//       broadcast values, E_ss = L_ss^-T (so that (b) is a product).  Pivot chain in registers:
//       dg -= l^2 -> shuffle -> rsqrt.approx + one cubic correction; the next column's last update via shuffle.
//   (b) all warps, DMMA: X = B * E_ss for the 96 rows outside the sub-block (rows below: the L panel; rows above: E rows)
//   (c) all warps, DMMA: rank-32 update T(r,c) -= X(r,:) X(c,:)' of the columns to the right, for (r >= c) [block being
//       factored] or (r < o+32) [E rows that have started]
#include "common.cuh"

namespace {

constexpr int NB = 128;
constexpr int LDT = 132;     // == 4 (mod 16): DMMA A/B fragments read straight from T / XP without bank conflicts
constexpr int SB = 32;
constexpr int XLD = 132;     // XP[k][row]
constexpr int ELD = 36;      // Eop[k][n]
constexpr int POTF2_THREADS = 256;
constexpr int POTF2_SMEM = (NB * LDT + SB * XLD + SB * ELD + SB + NB) * 8;

// 1/sqrt(d) for the pivots: MUFU seed (rsqrt.approx, ~2^-20 relative) + one cubically convergent correction, all
// FMA: 49 cycles of latency against 94 for the libdevice rsqrt() and 160 for sqrt() (tools/lat_bench.cu).
__device__ __forceinline__ double pivot_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d * y, y, 1.0);                 // 1 - d y^2
    return fma(y * e, fma(0.375, e, 0.5), y);             // y (1 + e/2 + 3 e^2/8)
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* M, long ld, long strideM, int k0, int nb, double* W, long strideW, int* info, int infoOff) {
    extern __shared__ __align__(16) double sm[];
    double* T = sm;                       // 128 x 132
    double* XP = sm + NB * LDT;           // 32 x 132: XP[k*XLD + row] = staged panel X(row, k)
    double* Eop = XP + SB * XLD;          // 32 x 36 : Eop[k*ELD + n] = E_ss(k, n)
    double* dinvE = Eop + SB * ELD;       // 128: diagonal of L^-1
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int fr = lane >> 2;             // fragment row / n index
    const int fk = lane & 3;              // fragment k index
    double* Mb = M + (long)blockIdx.x * strideM;
    double* Wb = W + (long)blockIdx.x * strideW + (long)(k0 / NB) * NB * NB;

    // ---- load the lower triangle (identity padding for a ragged block, zeros above the diagonal)
    if (nb == NB) {
        // whole 16-byte row pairs, all in flight at once; the strict upper part is cleared afterwards
        const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(T);
        for (int idx = tid; idx < NB * NB / 2; idx += POTF2_THREADS) {
            const int r = (idx & 63) * 2, c = idx >> 6;
            if (r + 1 >= c) {
                const double* src = Mb + (long)(k0 + r) + (long)(k0 + c) * ld;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(tbase + (uint32_t)((c * LDT + r) * 8)), "l"(src));
            }
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;\n" ::);
        __syncthreads();
        for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
            const int r = idx & (NB - 1), c = idx >> 7;
            if (r < c) T[c * LDT + r] = 0.0;
        }
    } else {
#pragma unroll 8
        for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
            const int r = idx & (NB - 1), c = idx >> 7;
            double v = 0.0;
            if (r >= c) {
                if (r < nb && c < nb) v = Mb[(long)(k0 + r) + (long)(k0 + c) * ld];
                else if (r == c) v = 1.0;
            }
            T[c * LDT + r] = v;
        }
    }
    __syncthreads();

    for (int s = 0; s < NB / SB; ++s) {
        const int o = s * SB;
        // ================= (a) 32x32 Cholesky + its inverse transpose, warp 0 =================
        if (warp == 0) {
            double a[SB], e[SB];
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                a[j] = T[(o + j) * LDT + o + lane];
                e[j] = (j == lane) ? 1.0 : 0.0;
            }
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
                else if (lane == j) l = d * rinv;
                const double x = e[j] * rinv;          // E(lane, j); zero for lane > j
                e[j] = x;
                dg = fma(-l, l, dg);
                if (lane >= j) T[(o + j) * LDT + o + lane] = l;
                if (j + 1 < SB) {
                    const double ln = __shfl_sync(0xffffffffu, l, j + 1);
                    a[j + 1] = fma(-l, ln, a[j + 1]);
                    e[j + 1] = fma(-x, ln, e[j + 1]);
                }
                __syncwarp();
                // column j read back as 16-byte pairs (shared-memory loads are the scarce resource here)
                if ((j + 2) & 1) {
                    const double lc = T[(o + j) * LDT + o + j + 2];
                    if (j + 2 < SB) { a[j + 2] = fma(-l, lc, a[j + 2]); e[j + 2] = fma(-x, lc, e[j + 2]); }
                }
#pragma unroll
                for (int c = (j + 3) & ~1; c + 1 < SB; c += 2) {
                    const double2 lc = *reinterpret_cast<const double2*>(&T[(o + j) * LDT + o + c]);
                    a[c] = fma(-l, lc.x, a[c]);
                    e[c] = fma(-x, lc.x, e[c]);
                    a[c + 1] = fma(-l, lc.y, a[c + 1]);
                    e[c + 1] = fma(-x, lc.y, e[c + 1]);
                }
            }
            // E_ss: strict upper part into T, diagonal into dinvE, operand copies for (b) and (c)
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                const double v = (j >= lane) ? e[j] : 0.0;
                Eop[lane * ELD + j] = v;
                XP[j * XLD + o + lane] = v;
                if (j > lane) T[(o + j) * LDT + o + lane] = v;
                if (j == lane) dinvE[o + lane] = v;
            }
        }
        __syncthreads();
        // ================= (b) X = B * E_ss for the rows outside the diagonal sub-block, DMMA =================
        {
            const int rbase = warp * 16;                         // this warp's 16 rows (two 8-row blocks)
            const bool isdiag = (rbase >= o) && (rbase < o + SB);    // 16 | 32: uniform per warp
            if (!isdiag) {
                double af[2][8];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int k4 = 0; k4 < 8; ++k4) af[mi][k4] = T[(o + k4 * 4 + fk) * LDT + rbase + mi * 8 + fr];
                double acc[2][4][2];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    double bf[4];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) bf[ni] = Eop[(k4 * 4 + fk) * ELD + ni * 8 + fr];
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi][k4], bf[ni]);
                }
                __syncwarp();
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int r = rbase + mi * 8 + fr, n = ni * 8 + 2 * fk + h;
                            T[(o + n) * LDT + r] = acc[mi][ni][h];
                            XP[n * XLD + r] = acc[mi][ni][h];
                        }
            }
        }
        __syncthreads();
        // ================= (c) rank-32 update of the columns to the right, DMMA =================
        const int c_first = o + SB;
        const int nnb = (NB - c_first) / 8;                       // 8-column blocks to the right: 12, 8, 4, 0
        if (nnb > 0) {
            const int rbase = warp * 16;
            double af[2][8];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) af[mi][k4] = XP[(k4 * 4 + fk) * XLD + rbase + mi * 8 + fr];
            for (int nb0 = 0; nb0 < nnb; nb0 += 4) {
                // 16 rows x 32 columns at a time; skip when every entry is in the not-yet-started E region
                const int cb = c_first + nb0 * 8;
                if (rbase + 15 < cb && rbase >= c_first) continue;
                double acc[2][4][2];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    double bf[4];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) bf[ni] = XP[(k4 * 4 + fk) * XLD + cb + ni * 8 + fr];
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi][k4], bf[ni]);
                }
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int r = rbase + mi * 8 + fr, c = cb + ni * 8 + 2 * fk + h;
                            if (r >= c || r < c_first) T[c * LDT + r] -= acc[mi][ni][h];
                        }
            }
        }
        __syncthreads();
    }

    // ---- write back L (lower triangle of the valid part)
#pragma unroll 8
    for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
        const int r = idx & (NB - 1), c = idx >> 7;
        if (r >= c && r < nb && c < nb) Mb[(long)(k0 + r) + (long)(k0 + c) * ld] = T[c * LDT + r];
    }
    // ---- W = inv(L) (128x128 column-major, unit padded): W(r,c) = E(c,r) = T(row c, col r) for r > c.
    //      Transposed through XP (33-stride staging) so that both the shared reads and the global writes are contiguous.
    double* stage = XP;                                            // 128 x 33 doubles fit (4224 <= 32*132)
    for (int q = 0; q < NB / SB; ++q) {
        __syncthreads();
        // stage[kk*33 + j] = T(row kk, col 32q + j)   (threads walk down kk: contiguous in T)
        for (int idx = tid; idx < NB * SB; idx += POTF2_THREADS) {
            const int kk = idx & (NB - 1), j = idx >> 7;
            stage[kk * 33 + j] = T[(q * SB + j) * LDT + kk];
        }
        __syncthreads();
        // W(r = 32q + j, c = kk): threads walk along j: contiguous in global
        for (int idx = tid; idx < NB * SB; idx += POTF2_THREADS) {
            const int j = idx & (SB - 1), kk = idx >> 5;
            const int r = q * SB + j, c = kk;
            double w = 0.0;
            if (r > c) w = stage[kk * 33 + j];
            else if (r == c) w = dinvE[r];
            Wb[r + c * NB] = w;
        }
    }
}

}  // namespace

void launch_potf2(double* M, long ld, long strideM, int batch, int k0, int nb, double* W, long strideW, int* info,
                  int infoOff, cudaStream_t st) {
    static bool attr_set[FGP_MAXDEV] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < FGP_MAXDEV && !attr_set[dev]) {
        cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM);
        attr_set[dev] = true;
    }
    FGP_COUNT_LAUNCH();
    potf2_inv_kernel<<<batch, POTF2_THREADS, POTF2_SMEM, st>>>(M, ld, strideM, k0, nb, W, strideW, info, infoOff);
}
