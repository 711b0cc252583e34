// TMA variant of the FP64 tensor-core NT GEMM over 128x128 tiles (gemm_dmma.cu):  C (-)= A * B'.
// Same tile regions, validity rules, warp tiling (8 MMA warps, 32x64 per warp, DMMA m8n8k4) and numerics -- the
// accumulation order over k is identical, so both kernels give the same bits -- but the operands are moved by the
// tensor memory accelerator instead of per-thread cp.async:
//   * ONE thread is the producer: in the middle of a chunk it waits for the stage read during the previous chunk to be
//     released, arms its "full" mbarrier with the byte count and issues two cp.async.bulk.tensor copies (A and B slab of
//     32 k-columns); all other threads only wait, multiply and release -- no copy instructions or address arithmetic
//     between their DMMAs (cp.async cost ~4 issue cycles per 16-byte piece, 16 pieces per thread and chunk).  (A ninth,
//     dedicated producer warp would cap the kernel at 168 registers -- registers are granted per four warps -- and spill);
//   * tensor maps are 3-D (rows, k, batch element) over the column-major trapezoid; the box is 132 rows x 32 k, so a
//     slab lands in shared memory as [k][132 rows]: with a row pitch of 132 = 4 (mod 16) doubles the 16 fragment words
//     of a half-warp (4 rows x 4 k) fall into 16 different 8-byte banks -- conflict-free LDS.64 straight from the layout
//     TMA writes, no permutation pass.  The 4 extra rows per slab are the price (3 % more L2 traffic);
//   * k beyond K is zero-filled by the map's bounds (the map's k extent is exactly K); rows that do not exist logically
//     are either outside the map (zero) or feed only output rows / columns that are never stored.
// Reference: same as gemm_dmma.cu (dpotrf / dtrsm behind base::chol and backsolve, R/3_training_SF.R:249,267,
// R/4_prediction_SF.R:9,28-29, R/5_simulation_SF.R:10,14).
#include "common.cuh"
#include <cuda.h>
#include <algorithm>
#include <atomic>
#include <mutex>

namespace {

constexpr int KC = 32;                      // k extent of one pipeline stage
constexpr int STAGES = 3;
constexpr int PITCH = 132;                  // rows per k-column of a slab (128 + 4: bank-conflict-free fragment loads)
constexpr int OPER = PITCH * KC;            // doubles per operand per stage
constexpr int STAGE = 2 * OPER;
constexpr int SMEM_BYTES = STAGES * STAGE * 8;          // 202752
constexpr int MMA_WARPS = 8;
constexpr int THREADS = MMA_WARPS * 32;
constexpr uint32_t STAGE_TX = 2u * OPER * 8u;           // bytes both copies of a stage deliver

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct Valid {
    int mn, lo, holeHi, n;
    __device__ __forceinline__ bool operator()(int r) const { return r >= mn && (r < lo || (r >= holeHi && r < n)); }
};

__device__ __forceinline__ void mbar_init(uint32_t addr, int count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared.b64 st, [%0];\n}\n" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared.b64 st, [%0], %1;\n}\n" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, int parity) {
    int ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n selp.b32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(addr), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, int parity) {
    // bounded spin: a protocol bug must trap, never hang the device
    for (int spin = 0; !mbar_try_wait(addr, parity); ++spin)
        if (spin > (1 << 26)) __trap();
}
// one 3-D box (rows, k, batch element) from global to shared memory, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int row, int k, int b) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(row), "r"(k), "r"(b)
                 : "memory");
}

struct TileId { int I, J; bool inTri; };
__device__ __forceinline__ TileId decode_tile(const GemmArgs& g, int t) {
    TileId r;
    const int nTri = g.triT * (g.triT + 1) / 2;
    r.inTri = t < nTri;
    if (r.inTri) {
        // triangle, packed lower, in bands of SB tile rows: a band occupies the index range of its rows in row-major
        // order, but is walked column by column (all rows of the band for column 0, then column 1, ...) and ends with
        // its own diagonal triangle.  The ~148 tiles in flight then touch SB row blocks + ~148/SB column blocks of the
        // operand panel (24 x 2 MB at K = 2048) instead of a whole tile row's worth (45 x 2 MB), and every column
        // block is read from HBM once per band instead of once per row.  Measured on the K = 2048 update of n = 8192
        // (ncu, dram__bytes_read): row-major 713 MB, bands of 24 / 16 / 12 / 8 rows 547 / 509 / 492 / 517 MB; the run
        // time does not move (the kernel is bound by the DMMA pipe), an evict_last L2 hint on the TMA loads changes
        // nothing either (510 MB at 16).
        constexpr int SB = 12;
        int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= t) ++i;
        while (i * (i + 1) / 2 > t) --i;
        const int r0 = (i / SB) * SB;
        const int h = min(SB, g.triT - r0);
        int u = t - r0 * (r0 + 1) / 2;
        int li, lj;
        if (u < r0 * h) { lj = u / h; li = r0 + u % h; }
        else {
            u -= r0 * h;         // the band's diagonal triangle, row-major
            int d = (int)((sqrtf(8.0f * (float)u + 1.0f) - 1.0f) * 0.5f);
            while ((d + 1) * (d + 2) / 2 <= u) ++d;
            while (d * (d + 1) / 2 > u) --d;
            li = r0 + d; lj = r0 + (u - d * (d + 1) / 2);
        }
        r.I = g.tri0 + li;
        r.J = g.tri0 + lj;
    } else {         // rectangle: two row ranges x one column range
        t -= nTri;
        const int nr = g.rectNR + g.rectNR1;
        const int ri = t % nr;
        r.I = ri < g.rectNR ? g.rectR0 + ri : g.rectR1 + (ri - g.rectNR);
        r.J = g.rectC0 + t / nr;
    }
    return r;
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_dmma_tma_kernel(GemmArgs g, int ntiles, int flat, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
    extern __shared__ __align__(128) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nchunks = (g.K + KC - 1) / KC;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t fullB = sbase + SMEM_BYTES;            // full[s]  at +8 s
    const uint32_t emptyB = fullB + 8 * STAGES;           // empty[s] at +8 s
    const bool producer = tid == 0;                        // one thread feeds the ring between its warp's DMMAs
    if (producer) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, MMA_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
    }
    __syncthreads();
    // programmatic dependent launch: everything above overlapped the previous kernel's tail
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    if (g.pdl == 2) asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    // Work items are (tile, batch element) pairs.  Default: one grid row per batch element, a CTA walks the tiles
    // blockIdx.x, blockIdx.x + gridDim.x, ... of its element.  flat: a 1-D grid of one CTA per SM walks ALL pairs,
    // element-major, so the operand ring also runs through matrix boundaries -- for small-n batches with hundreds of
    // elements (lock-step candidate scoring) every CTA would otherwise do a single short tile and pay the ring's fill
    // and drain each time.  With few items per CTA the static deal balances worse than many short CTAs (one-row y
    // tiles and diagonal tiles are cheaper: -4 % at n = 1024 x 60), so the launcher turns it on only for long queues.
    const int nitems = flat ? ntiles * g.batch : ntiles;
    int tIdx = blockIdx.x;
    if (tIdx >= nitems) return;
    const Valid rv{g.rowMin, g.rowLo, g.rowHoleHi, g.nrows};
    const Valid cv{g.colMin, g.colLo, g.colHoleHi, g.ncols};
    // warps (wm, wn): 4 along m (32 rows each) x 2 along n (64 cols each); same assignment as gemm_dmma_kernel
    const int wn = (warp >> 2) & 1;
    const int wm = (warp + wn) & 3;
    const int sgnmask = g.accumulate ? (int)0x80000000 : 0;   // C -= A B': flip the sign bit of the A fragments (integer op)
    const int r_lane = lane >> 2;        // row within an 8x8 DMMA tile
    const int c_lane = 2 * (lane & 3);   // first of the two columns this lane owns
    const int aoff = (lane & 3) * PITCH + wm * 32 + r_lane;
    const int boff = (lane & 3) * PITCH + wn * 64 + r_lane;

    // ---- prologue: the first STAGES-1 chunks of this CTA's sequence (they may already belong to its second tile)
    if (producer) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            const int off = s / nchunks, cn = s - off * nchunks;
            const int tgt = tIdx + off * (int)gridDim.x;
            if (tgt < nitems) {
                const int b1 = flat ? tgt / ntiles : (int)blockIdx.y;
                const TileId t1 = decode_tile(g, tgt - (flat ? b1 * ntiles : 0));
                mbar_expect_tx(fullB + 8 * s, STAGE_TX);
                const uint32_t dst = sbase + (uint32_t)(s * STAGE * 8);
                tma_load_3d(dst, &mapA, fullB + 8 * s, t1.I * FGP_TILE, cn * KC, b1);
                tma_load_3d(dst + OPER * 8, &mapB, fullB + 8 * s, (t1.J - g.bTileBase) * FGP_TILE, cn * KC, b1);
            }
        }
    }

    int gc = 0;   // global chunk counter of this CTA (ring position)
    for (; tIdx < nitems; tIdx += gridDim.x) {
        const int b = flat ? tIdx / ntiles : (int)blockIdx.y;
        const TileId tl = decode_tile(g, tIdx - (flat ? b * ntiles : 0));
        double* C = g.C + (long)b * g.strideC;
        const int rowBase = tl.I * FGP_TILE;
        const int colBase = (tl.J - g.bTileBase) * FGP_TILE;
        const int wrow = rowBase + wm * 32;
        const int wcol = colBase + wn * 64;
        double acc[4][8][2];
        unsigned rmask = 0, cmask = 0;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
            if (rv(wrow + mi * 8 + r_lane)) rmask |= 1u << mi;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            if (cv(wcol + ni * 8 + c_lane)) cmask |= 1u << (2 * ni);
            if (cv(wcol + ni * 8 + c_lane + 1)) cmask |= 1u << (2 * ni + 1);
        }
        // a warp is idle when none of its 32 rows / 64 columns exists (ragged edge tiles, the one-row y tile)
        bool active = __any_sync(0xffffffffu, rmask != 0) && __any_sync(0xffffffffu, cmask != 0);
        // diagonal tile of the symmetric update: warps entirely above the diagonal do nothing
        if (g.accumulate && tl.inTri && tl.I == tl.J && (wm * 32 + 31 < wn * 64)) active = false;
        if (!active) rmask = 0;
        const bool full = __all_sync(0xffffffffu, rmask == 0xFu && cmask == 0xFFFFu);
        unsigned mval = 0;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
            if (__any_sync(0xffffffffu, (rmask >> mi) & 1u)) mval |= 1u << mi;
        double* Cw = C + (long)(wrow + r_lane) + (long)(wcol + c_lane) * g.ldc;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                acc[mi][ni][0] = 0.0;
                acc[mi][ni][1] = 0.0;
            }
        if (g.accumulate && active) {
            if (full) {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        // streaming loads / stores: a C tile is touched once per launch and must not push the operand
                        // panel, which every tile row and column re-reads, out of L2
                        acc[mi][ni][0] = __ldcs(&Cw[mi * 8 + (long)(ni * 8) * g.ldc]);
                        acc[mi][ni][1] = __ldcs(&Cw[mi * 8 + (long)(ni * 8 + 1) * g.ldc]);
                    }
            } else {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const bool rok = (rmask >> mi) & 1u;
                        if (rok && ((cmask >> (2 * ni)) & 1u)) acc[mi][ni][0] = Cw[mi * 8 + (long)(ni * 8) * g.ldc];
                        if (rok && ((cmask >> (2 * ni + 1)) & 1u)) acc[mi][ni][1] = Cw[mi * 8 + (long)(ni * 8 + 1) * g.ldc];
                    }
            }
        }

        for (int c = 0; c < nchunks; ++c, ++gc) {
            const int s = gc % STAGES;
            mbar_wait(fullB + 8 * s, (gc / STAGES) & 1);                 // both slabs of chunk gc have landed
            // the chunk STAGES-1 ahead in this CTA's sequence: still this tile, or already one of its next tiles
            const int ahead = c + STAGES - 1;
            const int off = ahead / nchunks;
            const int cn = ahead - off * nchunks;
            const int tgt = tIdx + off * (int)gridDim.x;
            const bool doCopy = tgt < nitems;
            const int bn = (flat && doCopy) ? tgt / ntiles : b;               // batch element / tile of that chunk
            const int tn = tgt - (flat ? bn * ntiles : 0);
            if (off > 0 && doCopy && cn == 0 && g.accumulate) {
                // pull that tile's C into L2: 128 columns x 8 lines of 128 bytes, 4 lines per thread
                const TileId t1 = decode_tile(g, tn);
                const double* Cn = g.C + (long)bn * g.strideC + (long)t1.I * FGP_TILE + (long)((t1.J - g.bTileBase) * FGP_TILE) * g.ldc;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = tid + THREADS * i;
                    const int col = q >> 3, seg = q & 7;
                    if (rv(t1.I * FGP_TILE + seg * 16) && cv((t1.J - g.bTileBase) * FGP_TILE + col))
                        asm volatile("prefetch.global.L2 [%0];\n" ::"l"(Cn + seg * 16 + (long)col * g.ldc));
                }
            }
            // panel solve: B = inv(L_kk) is lower triangular, B(c, k) = 0 for k > c -> columns < 64 need k < 64 only
            const bool skipChunk = g.bLowerTri && (c * KC >= (wn + 1) * 64);
            const bool mult = active && !skipChunk;
            const double* As = smem + s * STAGE + aoff;
            const double* Bs = smem + s * STAGE + OPER + boff;
#pragma unroll
            for (int kk = 0; kk < KC / 4; ++kk) {
                if (kk == KC / 8 && producer && doCopy) {
                    // refill of the stage read during the previous chunk: every warp must have left it.  Issued in the
                    // middle of a chunk so that the warps may drift ~1.5 chunks apart instead of meeting at a barrier
                    const int sn = (gc + STAGES - 1) % STAGES;
                    if (gc + STAGES - 1 >= STAGES) mbar_wait(emptyB + 8 * sn, (((gc + STAGES - 1) / STAGES) - 1) & 1);
                    int rb = rowBase, cb = colBase;
                    if (off > 0) {
                        const TileId t1 = decode_tile(g, tn);
                        rb = t1.I * FGP_TILE; cb = (t1.J - g.bTileBase) * FGP_TILE;
                    }
                    mbar_expect_tx(fullB + 8 * sn, STAGE_TX);
                    const uint32_t dst = sbase + (uint32_t)(sn * STAGE * 8);
                    tma_load_3d(dst, &mapA, fullB + 8 * sn, rb, cn * KC, bn);
                    tma_load_3d(dst + OPER * 8, &mapB, fullB + 8 * sn, cb, cn * KC, bn);
                }
                if (mult) {
                    double a[4], bb[8];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const double v = As[kk * 4 * PITCH + mi * 8];
                        a[mi] = __hiloint2double(__double2hiint(v) ^ sgnmask, __double2loint(v));
                    }
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) bb[ni] = Bs[kk * 4 * PITCH + ni * 8];
                    if (mval == 0xFu) {
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                            for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
                    } else {
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
                            if ((mval >> mi) & 1u) {
#pragma unroll
                                for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
                            }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(emptyB + 8 * s);                  // this warp is done reading stage s
        }

        if (active) {
            if (full) {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        __stcs(&Cw[mi * 8 + (long)(ni * 8) * g.ldc], acc[mi][ni][0]);
                        __stcs(&Cw[mi * 8 + (long)(ni * 8 + 1) * g.ldc], acc[mi][ni][1]);
                    }
            } else {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const bool rok = (rmask >> mi) & 1u;
                        if (rok && ((cmask >> (2 * ni)) & 1u)) Cw[mi * 8 + (long)(ni * 8) * g.ldc] = acc[mi][ni][0];
                        if (rok && ((cmask >> (2 * ni + 1)) & 1u)) Cw[mi * 8 + (long)(ni * 8 + 1) * g.ldc] = acc[mi][ni][1];
                    }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

// 3-D map over a column-major operand: (rows, k, batch element); box = PITCH rows x KC k x 1
bool make_map(CUtensorMap* m, const double* base, long rows, long K, long ld, long batch, long strideElems) {
    cuuint64_t dims[3] = {(cuuint64_t)rows, (cuuint64_t)K, (cuuint64_t)std::max<long>(1, batch)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)(batch > 1 ? strideElems : (long)ld * K) * 8};
    cuuint32_t box[3] = {(cuuint32_t)PITCH, (cuuint32_t)KC, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (strides[1] == 0 || (strides[1] & 15)) strides[1] = (cuuint64_t)ld * 8 * (cuuint64_t)std::max<long>(1, K);
    return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// returns false when the TMA path cannot take this launch (driver entry point missing, operand not 16-byte aligned):
// the caller then uses the cp.async kernel
bool launch_gemm_tma(const GemmArgs& g, cudaStream_t st) {
    static std::atomic<bool> attr_set[FGP_MAXDEV];        // zero-initialised; setting the attribute twice is harmless
    static std::atomic<int> nsm[FGP_MAXDEV];
    // (several host threads -- one per GPU of a context, the worker contexts of fgp_fit_batch -- may arrive here at once)
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            g_encode = (EncodeTiledFn)fn;
        else cudaGetLastError();
    });
    if (!g_encode) return false;
    if (((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15) || (g.lda & 1) || (g.ldb & 1) || (g.strideA & 1) || (g.strideB & 1)) return false;
    const int ntiles = g.triT * (g.triT + 1) / 2 + (g.rectNR + g.rectNR1) * g.rectNC;
    if (ntiles <= 0 || g.batch <= 0) return true;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= FGP_MAXDEV) dev = 0;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(gemm_dmma_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 64) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        int count = 0;
        cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev);
        nsm[dev] = count;
        attr_set[dev] = true;
    }
    CUtensorMap mapA, mapB;
    // rows beyond the logical extent are outside the map (zero-filled); k beyond K likewise
    const long rowsA = g.nrows, rowsB = g.bLowerTri ? FGP_TILE : g.ncols;
    if (!make_map(&mapA, g.A, rowsA, g.K, g.lda, g.batch, g.strideA) || !make_map(&mapB, g.B, rowsB, g.K, g.ldb, g.batch, g.strideB)) return false;
    const int smCount = nsm[dev];
    const int sms = smCount > 0 ? smCount : 148;
    int tpc = std::max(1, std::min(16, 3072 / std::max(1, g.K)));       // tiles per CTA: K=128 -> 16 ... K>=2048 -> 1
    if (g.maxTilesPerCta > 0) tpc = std::min(tpc, g.maxTilesPerCta);
    int gx = ntiles;
    if (g.maxTilesPerCta != 1) {
        const int want = (ntiles + tpc - 1) / tpc;
        gx = want <= sms ? std::min(ntiles, sms) : std::min(ntiles, ((want + sms - 1) / sms) * sms);
    }
    // long queues of short tiles: one persistent CTA per SM over all (tile, element) pairs (see the kernel)
    int flat = 0;
    if (g.flat > 0 && g.batch > 1 && g.maxTilesPerCta != 1 && g.K <= 1024 && (long)ntiles * g.batch >= (long)g.flat * sms) {
        flat = 1;
        gx = sms;
    }
    FGP_COUNT_LAUNCH();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(gx, flat ? 1 : g.batch); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES + 64; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g.pdl ? 1 : 0;
    GemmArgs gg = g;
    if (gg.pdl == 2 && (long)gx * (flat ? 1 : g.batch) > sms) gg.pdl = 1;
    return cudaLaunchKernelEx(&cfg, gemm_dmma_tma_kernel, gg, ntiles, flat, mapA, mapB) == cudaSuccess;
}
