// FP64 tensor-core (DMMA m8n8k4) NT GEMM over 128x128 tiles:  C (-)= A * B'.
//
// This one kernel is the trailing update of the blocked Cholesky (SYRK/GEMM: C -= P_I P_J', reference
// base::chol -> dpotrf at R/3_training_SF.R:249, R/4_prediction_SF.R:28, R/5_simulation_SF.R:14), the panel
// triangular solve (X = P_I * inv(L_kk)', reference backsolve at R/4_prediction_SF.R:9,29) and the Schur
// complement of simulate (K.ss - t(LInvK) %*% LInvK, R/5_simulation_SF.R:14).
//
// sm_100a facts that shape it: tcgen05/TMEM have no f64 kind; every FP64 mma.sync shape lowers
// to DMMA.8x8x4, whose operands are register fragments (A: 1 double/lane, B: 1 double/lane, C: 2 doubles/lane).
// So: 8 warps, warp tile 32x64 (4x8 DMMA tiles -> 64 accumulator doubles per lane), operands staged through
// shared memory by cp.async in a [row-block of 8][k][8 rows] layout that makes every fragment load a
// conflict-free 256-byte LDS.64, 3-stage pipeline over K chunks of 32.
#include "common.cuh"
#include <atomic>
#include <algorithm>

#ifdef FGP_GEMM_DEBUG
#define DBG(bit) (g.dbg & (bit))
#else
#define DBG(bit) 0
#endif

namespace {

constexpr int KC = 32;                      // k extent of one pipeline stage
constexpr int STAGES = 3;
constexpr int OPER = 128 * KC;              // doubles per operand per stage (32 KB)
constexpr int STAGE = 2 * OPER;             // A + B
constexpr int GEMM_SMEM = STAGES * STAGE * 8;
constexpr int GEMM_THREADS = 256;

// all-or-nothing 16-byte copy: zero-fill when `ignore` (a variable src-size makes ptxas emit a long branchy
// sequence per copy -- measured: it halved the kernel's throughput)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ignore) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n cp.async.cg.shared.global [%0], [%1], 16, p;\n}\n" ::"r"(dst),
        "l"(src), "r"((int)ignore));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct Valid {
    int mn, lo, holeHi, n;
    __device__ __forceinline__ bool operator()(int r) const { return r >= mn && (r < lo || (r >= holeHi && r < n)); }
};

// Shared-memory operand layout, per stage: [row block of 8: 16][k group of 4: KC/4][32 doubles], the 32 doubles of
// one (8 rows x 4 k) fragment ordered [row>>2][k&3][row&3].  A DMMA fragment load (lane -> row lane>>2, k lane&3) then
// reads 32 distinct consecutive doubles split 16 + 16 between the half-warps: conflict-free LDS.64.
__device__ __forceinline__ int frag_lane_off(int lane) { return (lane >> 4) * 16 + (lane & 3) * 4 + ((lane >> 2) & 3); }

// Per-thread copy plan for one operand: 8 pieces of 16 bytes (2 rows x 1 k) per K chunk.
// piece i: row block rb = (tid>>7) + 2i, rows 2*(tid&3) + {0,1}, k = (tid>>2)&31
struct CopyPlan {
    const double* gptr;   // global address of (row0 + (tid>>7)*8 + 2*(tid&3), k = kq); + i*16 rows, + kc*ld per chunk
    uint32_t sdst;        // shared byte offset inside an operand slab for piece 0; + i*(2*8*32*8) bytes
    unsigned okmask;      // bit i: the row pair of piece i has at least one valid row
    int kq;
};
__device__ __forceinline__ CopyPlan make_plan(const double* g, long ld, int row0, const Valid& v, int tid) {
    CopyPlan p;
    const int r2 = tid & 3;
    p.kq = (tid >> 2) & 31;
    const int rloc = (tid >> 7) * 8 + 2 * r2;
    p.gptr = g + row0 + rloc + (long)p.kq * ld;
    const int r = 2 * r2;
    p.sdst = (uint32_t)(((((tid >> 7) * (KC / 4) + (p.kq >> 2)) * 32) + (r >> 2) * 16 + (p.kq & 3) * 4 + (r & 3)) * 8);
    p.okmask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = row0 + rloc + 16 * i;
        // (invalid, valid) / (valid, invalid) pairs are copied whole: the partner row exists in memory (ld is padded)
        // and only feeds an output row / column that is never stored
        if (v(row) || v(row + 1)) p.okmask |= 1u << i;
    }
    return p;
}
__device__ __forceinline__ void issue_piece(const CopyPlan& p, uint32_t slab, const double* src, bool kok, int i) {
    const bool ok = kok && ((p.okmask >> i) & 1u);
    cp_async16(slab + p.sdst + (uint32_t)(i * 2 * (KC / 4) * 32 * 8), ok ? (const void*)(src + 16 * i) : (const void*)p.gptr, !ok);
}
__device__ __forceinline__ void issue_copy(const CopyPlan& p, uint32_t slab, long ld, int k0, int K) {
    const bool kok = (k0 + p.kq) < K;
    const double* src = p.gptr + (long)k0 * ld;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool ok = kok && ((p.okmask >> i) & 1u);
        cp_async16(slab + p.sdst + (uint32_t)(i * 2 * (KC / 4) * 32 * 8), ok ? (const void*)(src + 16 * i) : (const void*)p.gptr, !ok);
    }
}

__device__ __forceinline__ void mbar_init(uint32_t addr, int count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared.b64 st, [%0];\n}\n" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive(uint32_t addr) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];\n" ::"r"(addr) : "memory");
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


struct TileId { int I, J; bool inTri; };
__device__ __forceinline__ TileId decode_tile(const GemmArgs& g, int t) {
    TileId r;
    const int nTri = g.triT * (g.triT + 1) / 2;
    r.inTri = t < nTri;
    if (r.inTri) {   // triangle: row-major packed lower
        int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= t) ++i;
        while (i * (i + 1) / 2 > t) --i;
        r.I = g.tri0 + i;
        r.J = g.tri0 + (t - i * (i + 1) / 2);
    } else {         // rectangle: two row ranges x one column range
        t -= nTri;
        const int nr = g.rectNR + g.rectNR1;
        const int ri = t % nr;
        r.I = ri < g.rectNR ? g.rectR0 + ri : g.rectR1 + (ri - g.rectNR);
        r.J = g.rectC0 + t / nr;
    }
    return r;
}

// Persistent over tiles: CTA x works on tiles x, x + gridDim.x, ... of its batch element.  The cp.async ring runs
// THROUGH tile boundaries -- while the last chunks of a tile are multiplied, the first chunks of the CTA's next tile
// are already being copied and its C tile is prefetched into L2 -- so a tile boundary costs the epilogue stores and
// one L2-latency load of C instead of a CTA launch, a cold pipeline fill and a DRAM-latency load of C
// (measured per-tile fixed cost of the one-tile-per-CTA version: 5.4 us, i.e. 25 % of a K = 128 tile).
// Synchronisation is by mbarriers, not __syncthreads: per ring stage a "full" barrier completed by every thread's
// asynchronous cp.async arrival and an "empty" barrier with one arrival per warp.  The refill of a stage is issued
// in the second half of a chunk, so warps may drift ~1.5 chunks apart instead of meeting at a CTA barrier every 32 k
// (a CTA barrier per chunk cost 6-14 %: all warps hit their fragment-load latency together and the DMMA pipe drains).
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_dmma_kernel(GemmArgs g, int ntiles) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // warps (wm, wn): 4 along m (32 rows each) x 2 along n (64 cols each).  Scheduler = warp % 4: the two warps of one
    // scheduler get different wn (the panel solve's triangular B halves the work of wn = 0: both schedulers stay
    // balanced) and the two warps with the same wm sit on different schedulers (a tile with few valid rows, e.g. the
    // one-row y tile, still spreads over two of them)
    const int wn = (warp >> 2) & 1;
    const int wm = (warp + wn) & 3;
    const long b = blockIdx.y;
    const double* A = g.A + b * g.strideA;
    const double* B = g.B + b * g.strideB;
    double* C = g.C + b * g.strideC;
    const Valid rv{g.rowMin, g.rowLo, g.rowHoleHi, g.nrows};
    const Valid cv{g.colMin, g.colLo, g.colHoleHi, g.ncols};
    const int nchunks = (g.K + KC - 1) / KC;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const int fragoff = frag_lane_off(lane);
    const int sgnmask = g.accumulate ? (int)0x80000000 : 0;   // C -= A B': flip the sign bit of the A fragments (integer op)
    const int r_lane = lane >> 2;        // row within an 8x8 DMMA tile
    const int c_lane = 2 * (lane & 3);   // first of the two columns this lane owns

    const uint32_t fullB = sbase + GEMM_SMEM;            // full[s]  at +8 s
    const uint32_t emptyB = fullB + 8 * STAGES;          // empty[s] at +8 s
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(fullB + 8 * s, GEMM_THREADS); mbar_init(emptyB + 8 * s, GEMM_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // programmatic dependent launch: everything above overlapped the tail of the previous kernel of the stream; its
    // results are visible from here on.  Every kernel of the chain executes the wait (ordering stays transitive).
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    if (g.pdl == 2) asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    int tIdx = blockIdx.x;
    if (tIdx >= ntiles) return;
    // ---- prologue of the first tile: STAGES-1 chunks in flight
    {
        const TileId t0 = decode_tile(g, tIdx);
        const CopyPlan pa = make_plan(A, g.lda, t0.I * FGP_TILE, rv, tid);
        const CopyPlan pb = make_plan(B, g.ldb, (t0.J - g.bTileBase) * FGP_TILE, cv, tid);
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            // chunk s of the global sequence: tile offset s / nchunks (only != 0 when K <= 32), chunk s % nchunks
            const int off = s / nchunks, cn = s - off * nchunks;
            const int tgt = tIdx + off * (int)gridDim.x;
            if (off == 0) {
                issue_copy(pa, sbase + (uint32_t)(s * STAGE * 8), g.lda, cn * KC, g.K);
                issue_copy(pb, sbase + (uint32_t)((s * STAGE + OPER) * 8), g.ldb, cn * KC, g.K);
            } else if (tgt < ntiles) {
                const TileId t1 = decode_tile(g, tgt);
                const CopyPlan qa = make_plan(A, g.lda, t1.I * FGP_TILE, rv, tid);
                const CopyPlan qb = make_plan(B, g.ldb, (t1.J - g.bTileBase) * FGP_TILE, cv, tid);
                issue_copy(qa, sbase + (uint32_t)(s * STAGE * 8), g.lda, cn * KC, g.K);
                issue_copy(qb, sbase + (uint32_t)((s * STAGE + OPER) * 8), g.ldb, cn * KC, g.K);
            }
            mbar_cp_async_arrive(fullB + 8 * s);
        }
    }

    int gc = 0;   // global chunk counter of this CTA (ring position)
    for (; tIdx < ntiles; tIdx += gridDim.x) {
        const TileId tl = decode_tile(g, tIdx);
        const int rowBase = tl.I * FGP_TILE;
        const int colBase = (tl.J - g.bTileBase) * FGP_TILE;
        const CopyPlan pa = make_plan(A, g.lda, rowBase, rv, tid);
        const CopyPlan pb = make_plan(B, g.ldb, colBase, cv, tid);

        // ---- accumulators: C tile (update) or zero (solve)
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
        // 8-row blocks of this warp that hold at least one valid row (warp-uniform): the others are not multiplied
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
        if (g.accumulate && active && !DBG(1)) {
            if (full) {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        acc[mi][ni][0] = Cw[mi * 8 + (long)(ni * 8) * g.ldc];
                        acc[mi][ni][1] = Cw[mi * 8 + (long)(ni * 8 + 1) * g.ldc];
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

        // ---- main loop over K chunks
        for (int c = 0; c < nchunks; ++c, ++gc) {
            mbar_wait(fullB + 8 * (gc % STAGES), (gc / STAGES) & 1);     // chunk gc has landed (all threads' pieces)
            // the chunk STAGES-1 ahead in this CTA's sequence: still this tile, or already one of its next tiles
            const int ahead = c + STAGES - 1;
            const int off = ahead / nchunks;                 // 0 = this tile (1 or 2 only near the end of the tile)
            const int cn = ahead - off * nchunks;
            const int tgt = tIdx + off * (int)gridDim.x;
            const bool doCopy = (tgt < ntiles) && !DBG(4);
            const uint32_t stA = sbase + (uint32_t)(((gc + STAGES - 1) % STAGES) * STAGE * 8);
            const uint32_t stB = stA + OPER * 8;
            CopyPlan qa = pa, qb = pb;
            if (off > 0 && doCopy) {
                const TileId t1 = decode_tile(g, tgt);
                qa = make_plan(A, g.lda, t1.I * FGP_TILE, rv, tid);
                qb = make_plan(B, g.ldb, (t1.J - g.bTileBase) * FGP_TILE, cv, tid);
                if (g.accumulate && cn == 0) {
                    // pull that tile's C into L2: 128 columns x 8 lines of 128 bytes, 4 lines per thread
                    const double* Cn = C + (long)t1.I * FGP_TILE + (long)((t1.J - g.bTileBase) * FGP_TILE) * g.ldc;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int q = tid + GEMM_THREADS * i;
                        const int col = q >> 3, seg = q & 7;
                        if (rv(t1.I * FGP_TILE + seg * 16) && cv((t1.J - g.bTileBase) * FGP_TILE + col))
                            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(Cn + seg * 16 + (long)col * g.ldc));
                    }
                }
            }
            const bool kokA = (cn * KC + qa.kq) < g.K, kokB = (cn * KC + qb.kq) < g.K;
            const double* srcA = qa.gptr + (long)(cn * KC) * g.lda;
            const double* srcB = qb.gptr + (long)(cn * KC) * g.ldb;
            // panel solve: B = inv(L_kk) is lower triangular, B(c, k) = 0 for k > c -> columns < 64 need k < 64 only
            const bool skipChunk = g.bLowerTri && (c * KC >= (wn + 1) * 64);
            if (active && !skipChunk && !DBG(16)) {
                const double* As = smem + (gc % STAGES) * STAGE + (wm * 4) * (KC / 4) * 32 + fragoff;
                const double* Bs = smem + (gc % STAGES) * STAGE + OPER + (wn * 8) * (KC / 4) * 32 + fragoff;
#pragma unroll
                for (int kk = 0; kk < KC / 4; ++kk) {
                    double a[4], bb[8];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const double v = As[(mi * (KC / 4) + kk) * 32];
                        a[mi] = __hiloint2double(__double2hiint(v) ^ sgnmask, __double2loint(v));
                    }
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) bb[ni] = Bs[(ni * (KC / 4) + kk) * 32];
                    if (doCopy && kk >= KC / 8) {
                        // refill of the stage read during the previous chunk: every warp must have left it
                        if (kk == KC / 8 && (gc + STAGES - 1) >= STAGES)
                            mbar_wait(emptyB + 8 * ((gc + STAGES - 1) % STAGES), (((gc + STAGES - 1) / STAGES) - 1) & 1);
                        issue_piece(qa, stA, srcA, kokA, 2 * (kk - KC / 8));
                        issue_piece(qb, stB, srcB, kokB, 2 * (kk - KC / 8));
                        issue_piece(qa, stA, srcA, kokA, 2 * (kk - KC / 8) + 1);
                        issue_piece(qb, stB, srcB, kokB, 2 * (kk - KC / 8) + 1);
                    }
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
            } else if (doCopy) {
                if ((gc + STAGES - 1) >= STAGES)
                    mbar_wait(emptyB + 8 * ((gc + STAGES - 1) % STAGES), (((gc + STAGES - 1) / STAGES) - 1) & 1);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    issue_piece(qa, stA, srcA, kokA, i);
                    issue_piece(qb, stB, srcB, kokB, i);
                }
            }
            if (doCopy) mbar_cp_async_arrive(fullB + 8 * ((gc + STAGES - 1) % STAGES));
            __syncwarp();
            if (lane == 0) mbar_arrive(emptyB + 8 * (gc % STAGES));     // this warp is done reading stage gc
        }

        // ---- epilogue
        if (active && !DBG(2)) {
            if (full) {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        Cw[mi * 8 + (long)(ni * 8) * g.ldc] = acc[mi][ni][0];
                        Cw[mi * 8 + (long)(ni * 8 + 1) * g.ldc] = acc[mi][ni][1];
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
    asm volatile("cp.async.wait_all;\n" ::: "memory");
}


}  // namespace

bool launch_gemm_tma(const GemmArgs& g, cudaStream_t st);

void launch_gemm(const GemmArgs& g, cudaStream_t st) {
    // TMA operand path (gemm_tma.cu): same numerics; falls back here when it cannot take the launch
    if (g.tma && launch_gemm_tma(g, st)) return;
    static std::atomic<bool> attr_set[FGP_MAXDEV];        // zero-initialised; several host threads may arrive at once
    static std::atomic<int> nsm[FGP_MAXDEV];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= FGP_MAXDEV) dev = 0;
    if (!attr_set[dev]) {
        cudaFuncSetAttribute(gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM + 64);
        int count = 0;
        cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev);
        nsm[dev] = count;
        attr_set[dev] = true;
    }
    const int ntiles = g.triT * (g.triT + 1) / 2 + (g.rectNR + g.rectNR1) * g.rectNC;
    if (ntiles <= 0 || g.batch <= 0) return;
    // semi-persistent: a CTA walks several tiles (operand prefetch across tile boundaries), but holds its SM for at
    // most ~0.4 ms so that other streams' panel kernels (diagonal block, panel solve) get an SM quickly
    const int smCount = nsm[dev];
    const int sms = smCount > 0 ? smCount : 148;
    int tpc = std::max(1, std::min(16, 3072 / std::max(1, g.K)));       // tiles per CTA: K=128 -> 16 ... K>=2048 -> 1
    if (g.maxTilesPerCta > 0) tpc = std::min(tpc, g.maxTilesPerCta);
    // whole waves: with more CTAs than SMs the grid is rounded up to a multiple of the SM count, otherwise the last,
    // partly filled wave still walks its full share of tiles while the other SMs idle (stand-alone, 1176 tiles, K = 512:
    // 196 CTAs x 6 tiles = 12 tile-times instead of 8, 20.9 instead of 30.9 TFLOP/s; inside a batch the other streams
    // fill those SMs and the step time does not change)
    int gx = ntiles;
    if (g.maxTilesPerCta != 1) {
        const int want = (ntiles + tpc - 1) / tpc;
        gx = want <= sms ? std::min(ntiles, sms) : std::min(ntiles, ((want + sms - 1) / sms) * sms);
    }
    if (g.dbg & 32) gx = ntiles;
    dim3 grid(gx, g.batch);
    FGP_COUNT_LAUNCH();
    if (g.pdl) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = GEMM_SMEM + 64; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        GemmArgs gg = g;
        if (gg.pdl == 2 && (long)gx * g.batch > sms) gg.pdl = 1;     // early release only when the whole grid is resident at once
        cudaLaunchKernelEx(&cfg, gemm_dmma_kernel, gg, ntiles);
        return;
    }
    gemm_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM + 64, st>>>(g, ntiles);     // + 2*STAGES mbarriers
}
