// The whole blocked Cholesky of a batch of (n + 1)-row trapezoids as ONE persistent dataflow kernel.
//
// Replaces, for the likelihood path (base::chol + backsolve of R/3_training_SF.R:249,267) and preMats
// (R/3_training_SF.R:262-267), the launch sequence of chol.cu -- per column tile: diagonal-block kernel, panel-solve
// GEMM, one or more trailing-update GEMMs -- by a single launch of one CTA per SM.  The unit of work is one 128x128
// tile (i, j) of one matrix, done left-looking by ONE CTA from start to finish:
//     acc  = A(i,j)                                      the assembled tile, read once
//     acc -= L(i, 0:j) L(j, 0:j)'                        K = 128 j through the TMA ring of gemm_tma.cu; a k-tile kt is
//                                                        requested as soon as L(i,kt) and L(j,kt) are flagged ready
//     i == j :  L(j,j), W_j = potf2(acc)                 potf2_body.cuh on the tile handed over in shared memory;
//               (L^-1 y)_j                               the y row rides along: the two warps of a diagonal tile that lie
//                                                        above the diagonal update it with FMAs in k order on the slabs
//                                                        passing through the ring, and it is solved against W_j here
//     i >  j :  L(i,j) = acc W_j'                        once W_j is flagged; DMMA from shared memory, zero blocks skipped
//     store, __threadfence, flag (i,j) / W_j / y_j ready
// Tiles are claimed from one atomic counter in an order in which every tile's inputs have smaller numbers -- column by
// column, inside a column the diagonal tiles of all matrices first, then row by row over the matrices -- so a CTA only
// ever waits for tiles that are already running or done on other resident CTAs (or claimed by a CTA that is finishing a
// tile with a smaller number): no deadlock whatever the number of resident CTAs.  Why: (1) the chain diagonal block ->
// solve -> next diagonal block is walked at flag latency (~1-2 us) instead of three kernel launches per column, and the
// CTAs holding later tiles fill in behind it; (2) every C tile is read and written once instead of once per panel
// level; (3) batches of small matrices need one launch per round; (4) no separate one-row tiles for y.
// Numerics: per entry the same sequence of fused multiply-adds in the same k order as the launch-per-step engine (whose
// updates also accumulate k ascending, merely interrupted by exact stores and loads; one FP64 DMMA m8n8k4 is four
// chained FMAs in k order, tools/dmma_order.cu), the same potf2 code; only products with exact zeros of W_j are left
// out: results are bit-identical (tests/test_gpu_parity.py::test_dataflow_cholesky_is_bit_identical).
// Scope: the plain (n + 1)-row trapezoid -- any n; rows and columns from n on are masked in the last tile row / column and
// the last diagonal block is padded with the identity as in potf2_inv_kernel -- no holes, nothing prefactored: predict /
// simulate / LOOCV / update trapezoids stay on chol.cu.
#include "potf2_body.cuh"
#include <cuda.h>
#include <algorithm>
#include <atomic>
#include <mutex>

namespace {

constexpr int KC = 32;
constexpr int STAGES = 3;
constexpr int PITCH = 132;
constexpr int OPER = PITCH * KC;
constexpr int STAGE = 2 * OPER;
constexpr int RING_BYTES = STAGES * STAGE * 8;          // 202752; the first two stages double as the 128 x 132 hand-over tile
constexpr int THREADS = 256;
constexpr int MMA_WARPS = 8;
constexpr uint32_t STAGE_TX = 2u * OPER * 8u;
constexpr uint32_t SLAB_TX = OPER * 8u;
constexpr int BAR_OFF = RING_BYTES;                     // full[3] | empty[3]
constexpr int WBAR_OFF = BAR_OFF + 64;                  // wfull[2]: the two W slabs of the solve
constexpr int PBAR_OFF = WBAR_OFF + 16;                 // potf2's 32 column barriers
constexpr int TASK_OFF = PBAR_OFF + 8 * potf2::SB;
constexpr int WMIN_OFF = TASK_OFF + 16;                 // per-warp minima of the ready scan (8 ints)
constexpr int AY_OFF = WMIN_OFF + 32;                   // the y row of the diagonal tile in flight (128 doubles)
constexpr int DAG_SMEM = AY_OFF + 8 * FGP_TILE;
static_assert(potf2::POTF2_WORK_DOUBLES * 8 <= RING_BYTES, "potf2 works inside the ring's memory");
static_assert(potf2::NB * potf2::LDT * 8 == 2 * STAGE * 8, "hand-over tile = stages 0 and 1");
static_assert(potf2::LDT == PITCH, "one layout for the hand-over tile, potf2's T and the solve's A operand");

#ifdef FGP_DAG_PROF
// development aid: SM clocks per phase summed over all CTAs: [0] claim + decode + ready scan  [1] tile loop (incl. waits for
// inputs)  [2] diagonal epilogue  [3] solve epilogue (incl. the wait for W_j)  [4] tasks  [5] diagonal tasks  [6] kernel
__device__ unsigned long long g_dag_prof[16];      // [8] solve: hand-over + W wait  [9] solve: products  [10] solve: stores + fence
#define DPROF(i, t0) do { if (threadIdx.x == 0) atomicAdd(&g_dag_prof[i], (unsigned long long)(clock64() - (t0))); } while (0)
#else
#define DPROF(i, t0) do { } while (0)
#endif

struct DagArgs {
    double* M; long ld; long strideM;
    int batch, T, n;               // T = n / 128 tile rows and columns; the y row (row n) rides along in the diagonal tasks
    int G;                         // tile columns per claim group (divides T)
    int matrixMajor;               // claim order below the diagonal tiles of a group: 1 = matrix, row; 0 = row, matrix
    int hasY;                      // 1: row n holds y' and is solved along (likelihood, preMats); 0: a plain n x n matrix
    // solve mode (predict / simulate: E tiles of extra rows from row rowOff on against the T prefactored column tiles)
    int rowOff, mExtra, E;
    int upper;                     // solve mode: the extra rows are identity rows (LOOCV): row tile r is zero left of column tile r
    double* W; long strideW;
    int* info;
    int* flagL;                    // [batch][T][T]  L(i,j) stored
    int* flagW;                    // [batch][T]     W_j stored
    int* flagY;                    // [batch][T]     (L^-1 y) of column tile j stored
    int* counter;
    int ntasks;
};

// Claim order of the tile tasks: task number -> (tile row i, tile column j, matrix b).  Groups of G tile columns; inside
// a group the G (G + 1) / 2 tiles of its triangle first (row by row, every matrix), then the tiles below -- matrix by
// matrix and row by row inside a matrix (matrixMajor: tiles that read the same L(j, :) are claimed back to back, start
// within microseconds of each other and share its chunks in L2), or row by row over the matrices.  G = 1 is plain
// column order: a column's diagonal tiles of all matrices, then the rest.  With G > 1 the tiles (i, j), (i, j+1), ... of
// one row run side by side on neighbouring CTAs and stream the same L(i, :) chunks at the same time, at the price of a
// short wait at their ends (tile (i, j+1) needs L(i, j)).  Whatever the parameters, every input of a task -- L(i, k) and
// L(j, k) for k < j, W_j, for diagonal tiles the y entries of the columns before -- belongs to a task with a SMALLER
// number: that is what makes the spin-waits of the kernel deadlock-free, and tests/test_dag_order_cpu.py checks it
// exhaustively through fgp_test_dag_order.
__host__ __device__ inline void dag_decode_task(int task, int T, int B, int G, int matrixMajor, int& i, int& j, int& b) {
    int qg = 0;
    long pre = 0;
    for (;;) {
        const long cnt = (long)B * (G * (G + 1) / 2 + (T - (qg + 1) * G) * G);
        if ((long)task < pre + cnt) break;
        pre += cnt; ++qg;
    }
    int local = (int)((long)task - pre);
    const int j0 = qg * G;
    int rr = 0;
    while (rr < G && local >= B * (rr + 1)) { local -= B * (rr + 1); ++rr; }
    if (rr < G) { i = j0 + rr; b = local / (rr + 1); j = j0 + local - b * (rr + 1); }
    else if (matrixMajor) {
        const int perB = (T - j0 - G) * G;
        b = local / perB;
        const int l2 = local - b * perB;
        i = j0 + G + l2 / G; j = j0 + l2 - (l2 / G) * G;
    } else {
        const int per = B * G;
        i = j0 + G + local / per;
        const int l2 = local - (local / per) * per;
        b = l2 / G; j = j0 + l2 - b * G;
    }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
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
// Waits in this kernel can legitimately last as long as the factorisation (a CTA holding a late tile waits for the
// chain), so the bound is a time, not a count: ~20 s of SM clock, then trap -- a protocol bug must never hang the device.
constexpr long long SPIN_LIMIT_CLOCKS = 40000000000LL;
__device__ __forceinline__ void mbar_wait(uint32_t addr, int parity) {
    if (mbar_try_wait(addr, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(addr, parity))
        if (clock64() - t0 > SPIN_LIMIT_CLOCKS) __trap();
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int row, int k, int b) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(row), "r"(k), "r"(b)
                 : "memory");
}
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// tile data were written with ordinary stores by another SM and are read here by the TMA unit (async proxy)
__device__ __forceinline__ void wait_flag(const int* p) {
    if (ld_acquire(p) == 0) {
        const long long t0 = clock64();
        while (ld_acquire(p) == 0)          // one thread per CTA polls: no back-off (it is worth 1-2 % on chain-bound factorisations)
            if (clock64() - t0 > SPIN_LIMIT_CLOCKS) __trap();
    }
    asm volatile("fence.proxy.async;\n" ::: "memory");
}

// The diagonal tile, handed over in shared memory: factor it, store W_j and flag it (the solves of the column start), then
// the y row of this column tile -- y_out(c) = sum_{k <= c} ay(k) W_j(c, k), the chain of FMAs in k order that the DMMA
// solve of the launch-per-step engine performs (one m8n8k4 DMMA IS four chained FMAs in k order, tools/dmma_order.cu) --
// then L(j,j).  Not inlined: the block algorithm runs at 255 registers on its own and must not share an allocation with
// the 128 accumulator registers of the tile loop.
__device__ __noinline__ void diag_epilogue(double* smem, uint32_t pbar, int* infoPtr, int infoOff, double* Mdiag, long ld, int nb, double* Wb,
                                           int* flagW, const double* ay, double* yOut, int* flagY) {
    potf2::factor_block(smem, pbar, infoPtr, infoOff);
    potf2::store_W(smem, Wb);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release(flagW, 1);
    if (yOut && threadIdx.x < nb) {
        const int c = threadIdx.x;
        const double* Tc = smem + c * potf2::LDT;                 // strict upper part of column c: E(k, c) = W(c, k), k < c
        const double* dinvE = smem + potf2::NB * potf2::LDT + potf2::SB * potf2::XLD + potf2::SB * potf2::ELD;
        double sacc = 0.0;
        for (int k = 0; k < c; ++k) sacc = fma(ay[k], Tc[k], sacc);
        sacc = fma(ay[c], dinvE[c], sacc);
        yOut[(long)c * ld] = sacc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release(flagY, 1);
    potf2::store_L(smem, Mdiag, ld, nb);
}

// MODE 0: factorisation; MODE 1: solve mode (extra rows against the prefactored columns).  A template parameter, not a field:
// the extra branches cost the tile loop of the factorisation registers it does not have (255, spills in the DMMA loop).
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
chol_dag_kernel(DagArgs g, const __grid_constant__ CUtensorMap mapM, const __grid_constant__ CUtensorMap mapW) {
    extern __shared__ __align__(128) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t fullB = sbase + BAR_OFF;
    const uint32_t emptyB = fullB + 8 * STAGES;
    const uint32_t wfullB = sbase + WBAR_OFF;
    const uint32_t pbar = sbase + PBAR_OFF;
    volatile int* sTask = reinterpret_cast<volatile int*>(reinterpret_cast<char*>(smem) + TASK_OFF);
    volatile int* sWarpMin = reinterpret_cast<volatile int*>(reinterpret_cast<char*>(smem) + WMIN_OFF);
    double* ay = reinterpret_cast<double*>(reinterpret_cast<char*>(smem) + AY_OFF);
    const bool producer = tid == 0;
    if (producer) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, MMA_WARPS); }
        mbar_init(wfullB, 1); mbar_init(wfullB + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&mapM)) : "memory");
        asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&mapW)) : "memory");
        sTask[0] = atomicAdd(g.counter, 1);
    }
    potf2::init_barriers(pbar);
    __syncthreads();

    // warps (wm, wn): 4 along m (32 rows each) x 2 along n (64 cols each), as in gemm_tma.cu
    const int wn = (warp >> 2) & 1;
    const int wm = (warp + wn) & 3;
    const int r_lane = lane >> 2;
    const int c_lane = 2 * (lane & 3);
    const int aoff = (lane & 3) * PITCH + wm * 32 + r_lane;
    const int boff = (lane & 3) * PITCH + wn * 64 + r_lane;
    const int B = g.batch, T = g.T;
    // With several matrices in flight a CTA claims its next tile while it finishes the current one (the atomic's round trip
    // hides under the epilogue).  With few, idle CTAs are waiting for exactly such a tile: claim only when free.
    const bool claimAhead = B >= 4;

    int gc = 0;   // chunks this CTA has sent through the ring so far (ring position and barrier phases)
    int task = sTask[0];
#ifdef FGP_DAG_PROF
    const long long tKernel = clock64();
#endif
    for (int it = 0; task < g.ntasks; ++it) {
#ifdef FGP_DAG_PROF
        long long tPh = clock64();
#endif
        int i, j, b;
        constexpr bool extra = MODE != 0;
        if (!extra) dag_decode_task(task, T, B, g.G, g.matrixMajor, i, j, b);
        else if (!g.upper) { j = task / g.E; i = T + (task - j * g.E); b = 0; }     // solve mode: column by column, the E row tiles of each
        else {
            // identity rows (X = L^-T is upper triangular): only the tiles (r, j >= r) exist -- column j has j + 1 of them
            j = (int)((sqrtf(8.0f * (float)task + 1.0f) - 1.0f) * 0.5f);
            while ((j + 1) * (j + 2) / 2 <= task) ++j;
            while (j * (j + 1) / 2 > task) --j;
            i = T + (task - j * (j + 1) / 2); b = 0;
        }
        // first k-tile of the accumulation: identity rows start at their own column tile
        const int cstart = (extra && g.upper) ? (i - T) * (FGP_TILE / KC) : 0;
        const bool diag = !extra && i == j;
        const int rowBase = extra ? g.rowOff + (i - T) * FGP_TILE : i * FGP_TILE, colBase = j * FGP_TILE;
        double* C = g.M + (long)b * g.strideM;
        // ready flags of this tile's own row; in solve mode the factor's tiles (and every W_j) are there from the start
        int* fLrow = extra ? g.flagL + (long)(i - T) * T : g.flagL + ((long)b * T + i) * T;
        const int* fLi = fLrow;
        const int* fLj = g.flagL + ((long)b * T + j) * T;
        const int* fY = g.flagY + (long)b * T;
        const int nchunks = j * (FGP_TILE / KC);
        // Which inputs of this tile are there already?  readyUpTo = the first k-tile with an input still missing (j: none
        // -- the usual case away from the start and the end of the factorisation).  Nobody polls for the k-tiles before
        // it; from there on the producer waits k-tile by k-tile as they are flagged.
        int firstMissing = j;
        for (int kt = (cstart >> 2) + tid; kt < j; kt += THREADS) {
            bool ok = ld_acquire(fLi + kt) != 0;
            if (!extra) {
                if (!diag) ok = ok && ld_acquire(fLj + kt) != 0;
                else if (g.hasY) ok = ok && ld_acquire(fY + kt) != 0;
            }
            if (!ok) firstMissing = min(firstMissing, kt);
        }
        firstMissing = __reduce_min_sync(0xffffffffu, firstMissing);
        if (lane == 0) sWarpMin[warp] = firstMissing;
        __syncthreads();                                    // (also: everybody has read the task slot)
        int readyUpTo = sWarpMin[0];
#pragma unroll
        for (int w = 1; w < MMA_WARPS; ++w) readyUpTo = min(readyUpTo, sWarpMin[w]);
#ifdef FGP_DAG_PROF
        DPROF(0, tPh); tPh = clock64();
        if (threadIdx.x == 0) { atomicAdd(&g_dag_prof[4], 1ULL); if (diag) atomicAdd(&g_dag_prof[5], 1ULL); }
#endif

        // diagonal tiles: the two warps that lie entirely above the diagonal carry the y row instead -- 64 columns each,
        // two per lane: ay(c) -= sum_k y(k) L(j-tile row c, k), FMAs in k order on the B slabs that pass through the ring
        const bool active = !(diag && (wm * 32 + 31 < wn * 64));
        const bool ywarp = diag && !active && g.hasY;
        const int ycol = (wm == 0 ? 0 : 64) + lane;
        const double* yrow = C + g.n;                                    // element k of the y row: yrow[k * ld]
        double ay0 = 0.0, ay1 = 0.0;
        // ragged last tile row / column (n not a multiple of 128): rows and columns from n on do not exist
        const int nbRows = extra ? min(FGP_TILE, g.mExtra - (i - T) * FGP_TILE) : min(FGP_TILE, g.n - rowBase);
        const int nbCols = min(FGP_TILE, g.n - colBase);
        const bool fullTile = nbRows == FGP_TILE && nbCols == FGP_TILE;
        if (ywarp) {
            if (ycol < nbCols) ay0 = __ldcg(yrow + (long)(colBase + ycol) * g.ld);
            if (ycol + 32 < nbCols) ay1 = __ldcg(yrow + (long)(colBase + ycol + 32) * g.ld);
        }
        double* Cw = C + (long)(rowBase + wm * 32 + r_lane) + (long)(colBase + wn * 64 + c_lane) * g.ld;

        // the producer's request for chunk cl of this tile (ring position G)
        auto issue = [&](int cl, int G) {
            if ((cl & 3) == 0 && (cl >> 2) >= readyUpTo) {
                wait_flag(fLi + (cl >> 2));
                if (!diag && !extra) wait_flag(fLj + (cl >> 2));
            }
            const int sn = G % STAGES;
            if (G >= STAGES) mbar_wait(emptyB + 8 * sn, ((G / STAGES) - 1) & 1);
            mbar_expect_tx(fullB + 8 * sn, STAGE_TX);
            const uint32_t dst = sbase + (uint32_t)(sn * STAGE * 8);
            tma_load_3d(dst, &mapM, fullB + 8 * sn, rowBase, cl * KC, b);
            tma_load_3d(dst + OPER * 8, &mapM, fullB + 8 * sn, colBase, cl * KC, b);
        };
        if (producer) {
            // the ring's memory was last touched by ordinary loads and stores (the previous tile's hand-over)
            asm volatile("fence.proxy.async;\n" ::: "memory");
#pragma unroll
            for (int s = 0; s < STAGES - 1; ++s)
                if (cstart + s < nchunks) issue(cstart + s, gc + s);
        }

        double acc[4][8][2];
        if (active && fullTile) {
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    acc[mi][ni][0] = __ldcs(&Cw[mi * 8 + (long)(ni * 8) * g.ld]);
                    acc[mi][ni][1] = __ldcs(&Cw[mi * 8 + (long)(ni * 8 + 1) * g.ld]);
                }
        } else {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
            if (active) {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const bool rok = wm * 32 + mi * 8 + r_lane < nbRows;
                        const int cl = wn * 64 + ni * 8 + c_lane;
                        if (rok && cl < nbCols) acc[mi][ni][0] = Cw[mi * 8 + (long)(ni * 8) * g.ld];
                        if (rok && cl + 1 < nbCols) acc[mi][ni][1] = Cw[mi * 8 + (long)(ni * 8 + 1) * g.ld];
                    }
            }
        }

        // ---- acc -= L(i, 0:j) L(j, 0:j)'
        for (int c = cstart; c < nchunks; ++c, ++gc) {
            const int s = gc % STAGES;
            double yk = 0.0;
            if (ywarp) {
                if ((c & 3) == 0 && (c >> 2) >= readyUpTo) {
                    const long long t0 = clock64();
                    while (ld_acquire(fY + (c >> 2)) == 0) {
                        __nanosleep(64);
                        if (clock64() - t0 > SPIN_LIMIT_CLOCKS) __trap();
                    }
                }
                yk = __ldcg(yrow + (long)(c * KC + lane) * g.ld);       // solved y entries of this chunk's 32 columns
            }
            mbar_wait(fullB + 8 * s, (gc / STAGES) & 1);
            const bool doCopy = c + STAGES - 1 < nchunks;
            const double* As = smem + s * STAGE + aoff;
            const double* Bs = smem + s * STAGE + OPER + boff;
            if (ywarp) {
                const double* Bc = smem + s * STAGE + OPER + ycol;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const double yv = -__shfl_sync(0xffffffffu, yk, k);
                    ay0 = fma(yv, Bc[k * PITCH], ay0);
                    ay1 = fma(yv, Bc[k * PITCH + 32], ay1);
                }
            }
#pragma unroll
            for (int kk = 0; kk < KC / 4; ++kk) {
                if (kk == KC / 8 && producer && doCopy) issue(c + STAGES - 1, gc + STAGES - 1);
                if (active) {
                    double a[4], bb[8];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const double v = As[kk * 4 * PITCH + mi * 8];
                        a[mi] = __hiloint2double(__double2hiint(v) ^ (int)0x80000000, __double2loint(v));   // C -= A B'
                    }
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) bb[ni] = Bs[kk * 4 * PITCH + ni * 8];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(emptyB + 8 * s);
        }
        __syncthreads();     // every warp has left the ring
#ifdef FGP_DAG_PROF
        DPROF(1, tPh); tPh = clock64();
#endif
        if (producer && claimAhead) sTask[(it + 1) & 1] = atomicAdd(g.counter, 1);

        // ---- hand the tile over in shared memory: [column][132 rows] -- potf2's T, and the solve's A operand ([k][row])
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = wm * 32 + mi * 8 + r_lane, c = wn * 64 + ni * 8 + c_lane + h;
                    double v = acc[mi][ni][h];
                    if (diag) {
                        // potf2's input: the lower triangle of the nb x nb block, identity padding, zeros above the diagonal
                        if (r < c) v = 0.0;
                        else if (r >= nbRows || c >= nbCols) v = (r == c) ? 1.0 : 0.0;
                    } else if (extra && c >= nbCols) v = 0.0;      // solve mode, ragged last column tile: k beyond n does not exist
                    smem[c * PITCH + r] = v;
                }

        if (diag) {
            if (ywarp) { ay[ycol] = ycol < nbCols ? ay0 : 0.0; ay[ycol + 32] = ycol + 32 < nbCols ? ay1 : 0.0; }
            __syncthreads();
            diag_epilogue(smem, pbar, g.info ? g.info + b : nullptr, colBase, C + (long)colBase + (long)colBase * g.ld, g.ld, nbCols,
                          g.W + (long)b * g.strideW + (long)j * FGP_TILE * FGP_TILE, g.flagW + (long)b * T + j, ay,
                          g.hasY ? C + g.n + (long)colBase * g.ld : nullptr, g.flagY + (long)b * T + j);
            __syncthreads();     // the hand-over area is free again
        } else {
            // ---- L(i,j) = acc W_j'   (W_j lower triangular: W(c, k) = 0 for k > c, so an 8-column block of the result
            //      needs k up to its last column only -- the zero products are skipped, block by block)
            const uint32_t wslab = sbase + (uint32_t)(2 * STAGE * 8);
            if (producer) {
                if (!extra) wait_flag(g.flagW + (long)b * T + j);
#pragma unroll
                for (int cw = 0; cw < 2; ++cw) {
                    mbar_expect_tx(wfullB + 8 * cw, SLAB_TX);
                    tma_load_3d(wslab + (uint32_t)(cw * OPER * 8), &mapW, wfullB + 8 * cw, 0, j * FGP_TILE + cw * KC, b);
                }
            }
            __syncthreads();     // hand-over tile complete
#ifdef FGP_DAG_PROF
            long long tS = clock64();
            DPROF(8, tPh);
#endif
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
#pragma unroll 1
            for (int cw = 0; cw < FGP_TILE / KC; ++cw) {
                mbar_wait(wfullB + 8 * (cw & 1), (cw >> 1) & 1);
                const double* As = smem + cw * KC * PITCH + aoff;
                const double* Bs = smem + 2 * STAGE + (cw & 1) * OPER + boff;
#pragma unroll
                for (int kk = 0; kk < KC / 4; ++kk) {
                    // result columns below k0 = cw*32 + kk*4 only meet zeros of W from here on
                    const int niSkip = (cw * KC + kk * 4 - wn * 64) >> 3;       // arithmetic shift: negative = nothing to skip
                    if (niSkip >= 8) continue;
                    double a[4], bb[8];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) a[mi] = As[kk * 4 * PITCH + mi * 8];
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) bb[ni] = Bs[kk * 4 * PITCH + ni * 8];
                    if (niSkip <= 0) {               // nothing to skip (most of the work): the plain unrolled block
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                            for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
                    } else {
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni)
                            if (ni >= niSkip) {
#pragma unroll
                                for (int mi = 0; mi < 4; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
                            }
                    }
                }
                __syncthreads();     // slab cw & 1 is free again
                if (producer && cw + 2 < FGP_TILE / KC) {
                    mbar_expect_tx(wfullB + 8 * (cw & 1), SLAB_TX);
                    tma_load_3d(wslab + (uint32_t)((cw & 1) * OPER * 8), &mapW, wfullB + 8 * (cw & 1), 0, j * FGP_TILE + (cw + 2) * KC, b);
                }
            }
#ifdef FGP_DAG_PROF
            DPROF(9, tS); tS = clock64();
#endif
            if (fullTile) {
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        Cw[mi * 8 + (long)(ni * 8) * g.ld] = acc[mi][ni][0];
                        Cw[mi * 8 + (long)(ni * 8 + 1) * g.ld] = acc[mi][ni][1];
                    }
            } else {         // ragged tile: rows / columns beyond the matrix do not exist
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
                        if (wm * 32 + mi * 8 + r_lane < nbRows) {
                            const int cl = wn * 64 + ni * 8 + c_lane;
                            if (!extra || cl < nbCols) Cw[mi * 8 + (long)(ni * 8) * g.ld] = acc[mi][ni][0];        // (factorisation:
                            if (!extra || cl + 1 < nbCols) Cw[mi * 8 + (long)(ni * 8 + 1) * g.ld] = acc[mi][ni][1];   //  all 128 columns exist)
                        }
            }
            __threadfence();
            __syncthreads();
#ifdef FGP_DAG_PROF
            DPROF(10, tS);
#endif
            if (producer) st_release(fLrow + j, 1);
        }
        if (producer && !claimAhead) sTask[(it + 1) & 1] = atomicAdd(g.counter, 1);
        __syncthreads();
#ifdef FGP_DAG_PROF
        DPROF(diag ? 2 : 3, tPh);
#endif
        task = sTask[(it + 1) & 1];
    }
#ifdef FGP_DAG_PROF
    DPROF(6, tKernel);
#endif
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

#ifdef FGP_DAG_PROF
extern "C" void fgp_dag_prof_read(unsigned long long* out8, int reset) {
    cudaMemcpyFromSymbol(out8, g_dag_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_dag_prof, z, sizeof(z)); }
}
#endif

// test support (CPU gate): the claim order of the dataflow kernel, so that its topological property can be checked without a GPU
extern "C" int fgp_test_dag_order(int T, int B, int G, int matrixMajor, int task, int* i, int* j, int* b) {
    if (T <= 0 || B <= 0 || G <= 0 || T % G != 0 || !i || !j || !b) return -1;
    const long ntasks = (long)B * T * (T + 1) / 2;
    if (task < 0 || task >= ntasks) return -1;
    dag_decode_task(task, T, B, G, matrixMajor, *i, *j, *b);
    return 0;
}

bool chol_dag_supported(int n, int nrows) { return n > 0 && (nrows == n + 1 || nrows == n); }

size_t chol_dag_flag_bytes(int n, int batch) {
    const size_t T = ((size_t)n + FGP_TILE - 1) / FGP_TILE;
    return sizeof(int) * ((size_t)batch * T * T + 2 * (size_t)batch * T + 4);
}
size_t chol_dag_solve_flag_bytes(int n, int mExtra) {
    const size_t T = ((size_t)n + FGP_TILE - 1) / FGP_TILE, E = ((size_t)mExtra + FGP_TILE - 1) / FGP_TILE;
    return sizeof(int) * (E * T + 4);
}

namespace {
bool dag_prepare(int* dev_out, int* sms_out) {
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
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= FGP_MAXDEV) dev = 0;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(chol_dag_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, DAG_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(chol_dag_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DAG_SMEM) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        int count = 0;
        cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev);
        nsm[dev] = count;
        attr_set[dev] = true;
    }
    *dev_out = dev;
    const int count = nsm[dev];
    *sms_out = count > 0 ? count : 148;
    return true;
}
}  // namespace

// Factors the `batch` matrices at M (n columns, leading dimension ld; hasY: n + 1 rows, the y row rides along) in place, W
// receives the inverse diagonal blocks, info the failing minor orders.  flags: chol_dag_flag_bytes() of device memory.
// Returns false when the launch cannot be taken (driver entry point missing, alignment) -- the caller then runs chol.cu.
bool launch_chol_dag(double* M, long ld, long strideM, int batch, int n, double* W, long strideW, int* info, void* flags,
                     cudaStream_t st, int group, int matrixMajor, int hasY) {
    if (n <= 0 || batch <= 0) return false;
    int dev, sms;
    if (!dag_prepare(&dev, &sms)) return false;
    if (((uintptr_t)M & 15) || ((uintptr_t)W & 15) || (ld & 1) || (strideM & 1) || (strideW & 1)) return false;
    DagArgs g{};
    g.M = M; g.ld = ld; g.strideM = strideM; g.batch = batch;
    g.T = (n + FGP_TILE - 1) / FGP_TILE; g.n = n;
    g.G = (group > 1 && g.T % group == 0) ? group : 1;
    g.matrixMajor = matrixMajor;
    g.hasY = hasY ? 1 : 0;
    g.W = W; g.strideW = strideW; g.info = info;
    const size_t nL = (size_t)batch * g.T * g.T, nW = (size_t)batch * g.T;
    g.flagL = (int*)flags; g.flagW = g.flagL + nL; g.flagY = g.flagW + nW; g.counter = g.flagY + nW;
    const long perMatrix = (long)g.T * (g.T + 1) / 2;
    if (perMatrix * batch > 0x7fffffffL / 2) return false;
    g.ntasks = (int)(perMatrix * batch);
    CUtensorMap mapM, mapW;
    if (!make_map(&mapM, M, n + (hasY ? 1 : 0), n, ld, batch, strideM)) return false;
    if (!make_map(&mapW, W, FGP_TILE, (long)g.T * FGP_TILE, FGP_TILE, batch, strideW)) return false;      // W_j unit-padded: always 128 x 128
    if (cudaMemsetAsync(flags, 0, chol_dag_flag_bytes(n, batch), st) != cudaSuccess) { cudaGetLastError(); return false; }
    FGP_COUNT_LAUNCH();
    chol_dag_kernel<0><<<std::min(sms, g.ntasks), THREADS, DAG_SMEM, st>>>(g, mapM, mapW);
    return cudaGetLastError() == cudaSuccess;
}

// Solve mode (predict / simulate, R/4_prediction_SF.R:9, R/5_simulation_SF.R:10): the mExtra rows from row rowOff on of M
// (n columns: K(new, train)) against the factor stored in rows 0..n-1 of the same matrix and its inverse diagonal blocks
// W -> L^-1 K(train, new) transposed, in place.  One tile task per (row tile of the block, column tile): it accumulates
// -= X(r, 0:j) L(j, 0:j)' as the tiles of its own row are flagged and multiplies by W_j'.  flags:
// chol_dag_solve_flag_bytes().  upper: the extra rows are the n rows of an identity matrix (LOOCV: X = L^-T, upper
// triangular -- only the tiles on and right of the diagonal are computed, each from its own column tile on).  Returns
// false when the launch cannot be taken.
bool launch_chol_dag_solve(double* M, long ld, int n, int rowOff, int mExtra, double* W, void* flags, cudaStream_t st, int upper) {
    if (n <= 0 || mExtra <= 0 || rowOff < n) return false;
    int dev, sms;
    if (!dag_prepare(&dev, &sms)) return false;
    if (((uintptr_t)M & 15) || ((uintptr_t)W & 15) || (ld & 1)) return false;
    DagArgs g{};
    g.M = M; g.ld = ld; g.strideM = 0; g.batch = 1;
    g.T = (n + FGP_TILE - 1) / FGP_TILE; g.n = n;
    g.G = 1; g.matrixMajor = 1; g.hasY = 0;
    g.rowOff = rowOff; g.mExtra = mExtra; g.E = (mExtra + FGP_TILE - 1) / FGP_TILE;
    g.W = W; g.strideW = 0; g.info = nullptr;
    g.flagL = (int*)flags; g.flagW = g.flagL; g.flagY = g.flagL; g.counter = g.flagL + (size_t)g.E * g.T;
    if ((long)g.T * g.E > 0x7fffffffL / 2) return false;
    g.upper = upper ? 1 : 0;
    if (upper && (g.E != g.T || mExtra != n)) return false;       // identity rows: one per column
    g.ntasks = upper ? g.T * (g.T + 1) / 2 : g.T * g.E;
    if (g.ntasks < 16) return false;      // a handful of tiles: the fixed costs of the persistent launch do not pay (measured: n = 1000, m = 10)
    CUtensorMap mapM, mapW;
    if (!make_map(&mapM, M, (long)rowOff + mExtra, n, ld, 1, 0)) return false;
    if (!make_map(&mapW, W, FGP_TILE, (long)g.T * FGP_TILE, FGP_TILE, 1, 0)) return false;
    if (cudaMemsetAsync(flags, 0, chol_dag_solve_flag_bytes(n, mExtra), st) != cudaSuccess) { cudaGetLastError(); return false; }
    FGP_COUNT_LAUNCH();
    chol_dag_kernel<1><<<std::min(sms, g.ntasks), THREADS, DAG_SMEM, st>>>(g, mapM, mapW);
    return cudaGetLastError() == cudaSuccess;
}
