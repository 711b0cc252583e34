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
// This header holds the block algorithm as three device functions -- load, factor, store -- over a shared-memory
// work area, so that potf2_inv_kernel (potf2.cu: one launch per diagonal block) and the persistent dataflow kernel
// (chol_dag.cu: the diagonal block arrives in registers) run the very same code on the same bits.
#pragma once
#include "common.cuh"

namespace potf2 {


constexpr int NB = 128;
constexpr int LDT = 132;     // == 4 (mod 16): DMMA A/B fragments read straight from T / XP without bank conflicts
constexpr int SB = 32;
constexpr int XLD = 132;     // XP[k][row]
constexpr int ELD = 36;      // Eop[k][n]
constexpr int POTF2_THREADS = 256;
// work area in doubles: T | XP | Eop | dinvE (NB + SB) | rinvS (SB); the SB mbarriers (8 bytes each) live where the caller says
constexpr int POTF2_WORK_DOUBLES = NB * LDT + SB * XLD + SB * ELD + NB + SB + SB;
constexpr int POTF2_SMEM = (POTF2_WORK_DOUBLES + SB) * 8;

// 1/sqrt(d) for the pivots: MUFU seed (rsqrt.approx, ~2^-20 relative) + one cubically convergent correction, all
// FMA: 49 cycles of latency against 94 for the libdevice rsqrt() and 160 for sqrt() (tools/lat_bench.cu).
__device__ __forceinline__ double pivot_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d * y, y, 1.0);                 // 1 - d y^2
    return fma(y * e, fma(0.375, e, 0.5), y);             // y (1 + e/2 + 3 e^2/8)
}

__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared.b64 st, [%0];\n}\n" ::"r"(addr));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, int parity) {
    int ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(addr), "r"(parity)
                 : "memory");
    return ok != 0;
}

// shared-memory accesses that keep their order among themselves but are no barrier for the arithmetic around them
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
    asm volatile("st.volatile.shared.f64 [%0], %1;\n" ::"r"(addr), "d"(v));
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_v2f64(uint32_t addr) {
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

#ifdef FGP_POTF2_PROF
__device__ long long g_potf2_prof[32];
#define PROF(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_potf2_prof[i] = clock64(); } while (0)
#else
#define PROF(i) do { } while (0)
#endif


// the SB mbarriers at barAddr: one per column of a sub-block (32 arrivals: the lanes of the chain warp), one phase per
// sub-step -- four phases per block, so a barrier set is reusable for any number of blocks without re-initialisation
__device__ __forceinline__ void init_barriers(uint32_t barAddr) {
    if (threadIdx.x < SB) {
        asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(barAddr + 8 * threadIdx.x), "r"(32));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
}

// T <- lower triangle of the nb x nb block at Mdiag (identity padding for a ragged block, zeros above the diagonal)
__device__ __forceinline__ void load_block(double* sm, const double* Mdiag, long ld, int nb) {
    double* T = sm;
    const int tid = threadIdx.x;
    const double* Mb = Mdiag;
    const int k0 = 0;
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

}

// factor the block in T in place: on return the lower triangle holds L, the strict upper triangle E = L^-T and dinvE the
// diagonal of L^-1.  infoPtr (may be null): receives infoOff + (order of the first non-positive minor) if it is still 0.
__device__ __forceinline__ void factor_block(double* sm, uint32_t barAddr, int* infoPtr, int infoOff) {
    double* T = sm;                       // 128 x 132
    double* XP = sm + NB * LDT;           // 32 x 132: XP[k*XLD + row] = staged panel X(row, k)
    double* Eop = XP + SB * XLD;          // 32 x 36 : Eop[k*ELD + n] = E_ss(k, n)
    double* dinvE = Eop + SB * ELD;       // 128: diagonal of L^-1
    double* rinvS = dinvE + NB + SB;      // 32: 1 / l_jj of the current sub-block (warp 0 -> warp 1)
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int fr = lane >> 2;             // fragment row / n index
    const int fk = lane & 3;              // fragment k index
    // Pipelined over the four sub-steps: the chain of sub-step s runs on the two warps that own its rows (2s, 2s+1)
    // WHILE the other six warps are still in the rank-32 update of sub-step s-1 -- those two warps have a single
    // update task each (the 16 x 32 halves of the next diagonal block), do it first, meet at a 64-thread barrier and
    // start the chain.  The inverse warp keeps E_ss in registers until the update has finished reading XP.
    double e[SB];                         // E_ss row of the inverse warp, written out after barrier C
    for (int s = 0; s < NB / SB; ++s) {
        const int o = s * SB;
        PROF(1 + 3 * s);
        // ================= (c) rank-32 update with the panel of sub-step s-1, DMMA =================
        if (s > 0) {
            const int c_first = o;                                    // columns right of sub-block s-1
            const int nnb = (NB - c_first) / 8;                       // 8-column blocks to the right: 12, 8, 4
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
        // ================= (a) 32x32 Cholesky (warp 2s) and its inverse transpose (warp 2s+1, one column behind) =====
        if ((warp >> 1) == s) {
            // both halves of the diagonal sub-block have their update (each of the two warps did one half)
            if (s > 0) asm volatile("bar.sync 1, 64;\n" ::: "memory");
            if ((warp & 1) == 0) {
                // No branch, barrier or fence inside the 32 steps: shared memory is touched through volatile asm only,
                // so that ptxas is free to interleave the trailing updates of step j-1 with the pivot chain of step j.
                double a[SB];
#pragma unroll
                for (int j = 0; j < SB; ++j) a[j] = T[(o + j) * LDT + o + lane];
                double dg = T[(o + lane) * LDT + o + lane];
                int firstBad = 0;
                const uint32_t tS = (uint32_t)__cvta_generic_to_shared(T + o * LDT + o);   // sub-block origin
                const uint32_t rS = (uint32_t)__cvta_generic_to_shared(rinvS);
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    double d = __shfl_sync(0xffffffffu, dg, j);
                    const bool bad = !(d > 0.0);   // not positive definite: remember the minor's order, keep going finite
                    firstBad = (bad && firstBad == 0) ? j + 1 : firstBad;
                    d = bad ? 1.0 : d;
                    const double rinv = pivot_rsqrt(d);
                    double l = a[j] * rinv;
                    l = lane > j ? l : (lane == j ? d * rinv : 0.0);
                    dg = fma(-l, l, dg);
                    if (lane >= j) sts_f64(tS + (uint32_t)((j * LDT + lane) * 8), l);
                    if (lane == 0) sts_f64(rS + j * 8, rinv);
                    if (j + 1 < SB) {
                        const double ln = __shfl_sync(0xffffffffu, l, j + 1);
                        a[j + 1] = fma(-l, ln, a[j + 1]);
                    }
                    // column j and 1/l_jj are in shared memory: every lane releases its own stores to the inverse warp
                    mbar_arrive(barAddr + 8 * j);
                    // column j read back as 16-byte pairs (shared-memory loads are the scarce resource here)
                    if (((j + 2) & 1) && j + 2 < SB) a[j + 2] = fma(-l, lds_f64(tS + (uint32_t)((j * LDT + j + 2) * 8)), a[j + 2]);
#pragma unroll
                    for (int c = (j + 3) & ~1; c + 1 < SB; c += 2) {
                        const double2 lc = lds_v2f64(tS + (uint32_t)((j * LDT + c) * 8));
                        a[c] = fma(-l, lc.x, a[c]);
                        a[c + 1] = fma(-l, lc.y, a[c + 1]);
                    }
                }
                if (firstBad && lane == 0 && infoPtr) atomicCAS(infoPtr, 0, infoOff + o + firstBad);   // the first failing minor wins
            } else {
                // E_ss = L_ss^-T, lane = row of E, with the columns the chain warp has released
#pragma unroll
                for (int j = 0; j < SB; ++j) e[j] = (j == lane) ? 1.0 : 0.0;
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    for (int spin = 0; !mbar_try_wait(barAddr + 8 * j, s & 1); ++spin)
                        if (spin > (1 << 24)) __trap();   // a protocol bug must trap, never hang the device
                    const double x = e[j] * rinvS[j];     // E(lane, j); zero for lane > j
                    e[j] = x;
                    if ((j + 1) & 1) {
                        const double lc = T[(o + j) * LDT + o + j + 1];
                        if (j + 1 < SB) e[j + 1] = fma(-x, lc, e[j + 1]);
                    }
#pragma unroll
                    for (int c = (j + 2) & ~1; c + 1 < SB; c += 2) {
                        const double2 lc = *reinterpret_cast<const double2*>(&T[(o + j) * LDT + o + c]);
                        e[c] = fma(-x, lc.x, e[c]);
                        e[c + 1] = fma(-x, lc.y, e[c + 1]);
                    }
                }
            }
        }
        __syncthreads();     // C: update s-1 and chain s are done; XP and Eop are free
        if (warp == 2 * s + 1) {
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
        __syncthreads();     // A
        PROF(2 + 3 * s);
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
        __syncthreads();     // B
        PROF(3 + 3 * s);
    }

}

// W = inv(L) -> Wb (128 x 128 column-major, unit padded): what the rest of the factorisation waits for, so it goes first
__device__ __forceinline__ void store_W(double* sm, double* Wb) {
    double* T = sm;
    double* dinvE = sm + NB * LDT + SB * XLD + SB * ELD;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    PROF(13);
    // ---- W = inv(L) first: it is what the rest of the factorisation waits for
    // ---- (128x128 column-major, unit padded): W(r,c) = E(c,r) = T(row c, col r) for r > c, i.e. T read
    //      along its leading dimension.  A half-warp takes 4 consecutive r x 4 consecutive c: with LDT = 4 (mod 16) the
    //      sixteen 8-byte words fall into sixteen different bank pairs, and every 4 r of one c are one 32-byte sector.
    for (int it = 0; it < (NB * NB) / (POTF2_THREADS * 1); ++it) {
        const int blk = it * 8 + warp;                      // 32 r-groups x 16 c-groups: 512 blocks of 4 r x 8 c
        const int r = (blk & 31) * 4 + (lane & 3);
        const int c = (blk >> 5) * 8 + (lane >> 2);
        double w = 0.0;
        if (r > c) w = T[r * LDT + c];
        else if (r == c) w = dinvE[r];
        Wb[r + c * NB] = w;
    }
    PROF(14);
}

// L -> the lower triangle of the block at Mdiag
__device__ __forceinline__ void store_L(double* sm, double* Mdiag, long ld, int nb) {
    double* T = sm;
    const int tid = threadIdx.x;
    double* Mb = Mdiag;
    const int k0 = 0;
    // ---- write back L (lower triangle of the valid part)
#pragma unroll 8
    for (int idx = tid; idx < NB * NB; idx += POTF2_THREADS) {
        const int r = idx & (NB - 1), c = idx >> 7;
        if (r >= c && r < nb && c < nb) Mb[(long)(k0 + r) + (long)(k0 + c) * ld] = T[c * LDT + r];
    }
    PROF(15);
}

__device__ __forceinline__ void store_block(double* sm, double* Mdiag, long ld, int nb, double* Wb) {
    store_W(sm, Wb);
    store_L(sm, Mdiag, ld, nb);
}

}  // namespace potf2
