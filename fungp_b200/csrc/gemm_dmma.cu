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

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_dmma_kernel(GemmArgs g) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm = warp & 3;   // 4 warps along m (32 rows each)
    const int wn = warp >> 2;  // 2 warps along n (64 cols each)

    // ---- tile decode: triangle (row-major packed lower) then rectangle
    int t = blockIdx.x, I, J;
    const int nTri = g.triT * (g.triT + 1) / 2;
    const bool inTri = t < nTri;
    if (inTri) {
        int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((i + 1) * (i + 2) / 2 <= t) ++i;
        while (i * (i + 1) / 2 > t) --i;
        I = g.tri0 + i;
        J = g.tri0 + (t - i * (i + 1) / 2);
    } else {
        t -= nTri;
        const int nr = g.rectNR + g.rectNR1;
        const int ri = t % nr;
        I = ri < g.rectNR ? g.rectR0 + ri : g.rectR1 + (ri - g.rectNR);
        J = g.rectC0 + t / nr;
    }
    const long b = blockIdx.y;
    const double* A = g.A + b * g.strideA;
    const double* B = g.B + b * g.strideB;
    double* C = g.C + b * g.strideC;

    const int rowBase = I * FGP_TILE;
    const int colBase = (J - g.bTileBase) * FGP_TILE;
    const Valid rv{g.rowMin, g.rowLo, g.rowHoleHi, g.nrows};
    const Valid cv{g.colMin, g.colLo, g.colHoleHi, g.ncols};
    const int nchunks = (g.K + KC - 1) / KC;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);

    const CopyPlan pa = make_plan(A, g.lda, rowBase, rv, tid);
    const CopyPlan pb = make_plan(B, g.ldb, colBase, cv, tid);

    // ---- prologue: first STAGES-1 chunks in flight
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nchunks) {
            issue_copy(pa, sbase + (uint32_t)(s * STAGE * 8), g.lda, s * KC, g.K);
            issue_copy(pb, sbase + (uint32_t)((s * STAGE + OPER) * 8), g.ldb, s * KC, g.K);
        }
        cp_async_commit();
    }

    // ---- accumulators: C tile (update) or zero (solve)
    const int r_lane = lane >> 2;        // row within an 8x8 DMMA tile
    const int c_lane = 2 * (lane & 3);   // first of the two columns this lane owns
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
    if (g.accumulate && inTri && I == J && (wm * 32 + 31 < wn * 64)) active = false;
    if (!active) rmask = 0;
    const bool full = __all_sync(0xffffffffu, rmask == 0xFu && cmask == 0xFFFFu);
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
    const int fragoff = frag_lane_off(lane);
    const int sgnmask = g.accumulate ? (int)0x80000000 : 0;   // C -= A B': flip the sign bit of the A fragments (integer op)
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<STAGES - 2>();
        if (!DBG(8)) __syncthreads();
        // the copies of chunk c+STAGES-1 are spread over the 8 k-steps below (one 16-byte piece of A and of B per
        // step and thread) instead of being issued in one burst after the barrier: a burst keeps every warp in the
        // LSU queue while the DMMA pipe drains (measured: -15 % at K = 512)
        const int cn = c + STAGES - 1;
        const bool doCopy = (cn < nchunks) && !DBG(4);
        const uint32_t stA = sbase + (uint32_t)((cn % STAGES) * STAGE * 8);
        const uint32_t stB = stA + OPER * 8;
        const bool kokA = (cn * KC + pa.kq) < g.K, kokB = (cn * KC + pb.kq) < g.K;
        const double* srcA = pa.gptr + (long)(cn * KC) * g.lda;
        const double* srcB = pb.gptr + (long)(cn * KC) * g.ldb;
        if (active && !DBG(16)) {
            const double* As = smem + (c % STAGES) * STAGE + (wm * 4) * (KC / 4) * 32 + fragoff;
            const double* Bs = smem + (c % STAGES) * STAGE + OPER + (wn * 8) * (KC / 4) * 32 + fragoff;
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
                if (doCopy) {
                    issue_piece(pa, stA, srcA, kokA, kk);
                    issue_piece(pb, stB, srcB, kokB, kk);
                }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
            }
        } else if (doCopy) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                issue_piece(pa, stA, srcA, kokA, i);
                issue_piece(pb, stB, srcB, kokB, i);
            }
        }
        cp_async_commit();
    }
    cp_async_wait<0>();

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

}  // namespace

void launch_gemm(const GemmArgs& g, cudaStream_t st) {
    static bool attr_set[FGP_MAXDEV] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < FGP_MAXDEV && !attr_set[dev]) {
        cudaFuncSetAttribute(gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
        attr_set[dev] = true;
    }
    const int ntiles = g.triT * (g.triT + 1) / 2 + (g.rectNR + g.rectNR1) * g.rectNC;
    if (ntiles <= 0 || g.batch <= 0) return;
    dim3 grid(ntiles, g.batch);
    FGP_COUNT_LAUNCH();
    gemm_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
}
