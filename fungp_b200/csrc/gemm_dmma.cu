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

namespace {

constexpr int KC = 32;                      // k extent of one pipeline stage
constexpr int STAGES = 3;
constexpr int OPER = 128 * KC;              // doubles per operand per stage (32 KB)
constexpr int STAGE = 2 * OPER;             // A + B
constexpr int GEMM_SMEM = STAGES * STAGE * 8;
constexpr int GEMM_THREADS = 256;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
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
    // bytes to copy for the even-aligned pair (r, r+1).  (invalid, valid) copies both: the first element exists in
    // memory (r+1 does) and only feeds an output row/column that is never stored.
    __device__ __forceinline__ int pair_bytes(int r) const { return (*this)(r + 1) ? 16 : ((*this)(r) ? 8 : 0); }
};

// stage one 128 x KC operand chunk: global (col-major, rows contiguous) -> smem [rb 16][k KC][8]
__device__ __forceinline__ void load_operand(double* sm, const double* g, long ld, int row0, int k0, int K,
                                             const Valid& v, int tid) {
    const int r2 = tid & 3;
    const int kq = (tid >> 2) & 31;
    const int k = k0 + kq;
    const bool kok = k < K;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rb = (tid >> 7) + 2 * i;
        const int row = row0 + rb * 8 + 2 * r2;
        int bytes = kok ? v.pair_bytes(row) : 0;
        const double* src = bytes ? (g + row + (long)k * ld) : g;
        cp_async16(sbase + (uint32_t)(((rb * KC + kq) * 8 + 2 * r2) * 8), src, bytes);
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

    // ---- prologue: first STAGES-1 chunks in flight
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nchunks) {
            load_operand(smem + s * STAGE, A, g.lda, rowBase, s * KC, g.K, rv, tid);
            load_operand(smem + s * STAGE + OPER, B, g.ldb, colBase, s * KC, g.K, cv, tid);
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
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            acc[mi][ni][0] = 0.0;
            acc[mi][ni][1] = 0.0;
        }
    if (g.accumulate && active) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const long r = wrow + mi * 8 + r_lane;
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                const long c = wcol + ni * 8 + c_lane;
                if ((rmask >> mi) & 1u) {
                    if ((cmask >> (2 * ni)) & 1u) acc[mi][ni][0] = C[r + c * g.ldc];
                    if ((cmask >> (2 * ni + 1)) & 1u) acc[mi][ni][1] = C[r + (c + 1) * g.ldc];
                }
            }
        }
    }

    // ---- main loop over K chunks
    const int fragoff = (lane & 3) * 8 + (lane >> 2);   // (k%4, row%8) position inside a [k][8] slab
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int cn = c + STAGES - 1;
            if (cn < nchunks) {
                double* st = smem + (cn % STAGES) * STAGE;
                load_operand(st, A, g.lda, rowBase, cn * KC, g.K, rv, tid);
                load_operand(st + OPER, B, g.ldb, colBase, cn * KC, g.K, cv, tid);
            }
            cp_async_commit();
        }
        if (active) {
            const double* As = smem + (c % STAGES) * STAGE + (wm * 4) * KC * 8 + fragoff;
            const double* Bs = smem + (c % STAGES) * STAGE + OPER + (wn * 8) * KC * 8 + fragoff;
#pragma unroll
            for (int kk = 0; kk < KC / 4; ++kk) {
                double a[4], bb[8];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = As[mi * KC * 8 + kk * 32];
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) bb[ni] = Bs[ni * KC * 8 + kk * 32];
                if (g.accumulate) {
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) a[mi] = -a[mi];
                }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bb[ni]);
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue
    if (active) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const long r = wrow + mi * 8 + r_lane;
            if (!((rmask >> mi) & 1u)) continue;
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                const long c = wcol + ni * 8 + c_lane;
                if ((cmask >> (2 * ni)) & 1u) C[r + c * g.ldc] = acc[mi][ni][0];
                if ((cmask >> (2 * ni + 1)) & 1u) C[r + (c + 1) * g.ldc] = acc[mi][ni][1];
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
    gemm_dmma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(g);
}
