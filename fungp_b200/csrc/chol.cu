// Host-side driver of the "trapezoid" blocked Cholesky and the small reductions around it.
//
// One engine serves every dense step of the reference's hot path:
//   fit       chol(R + nugget I), backsolve(t(U), y), sum(log(diag(U)))   R/3_training_SF.R:246-269
//   preMats   t(chol(sig2 (R + nugget I))), L^-1 y                       R/4_prediction_SF.R:24-32
//   predict   backsolve(L, K.tp)                                         R/4_prediction_SF.R:9
//   simulate  backsolve, K.ss - t(LInvK) LInvK, chol                      R/5_simulation_SF.R:10-14
//   LOOCV     solve(R) through L^-T                                       R/8_outilsStats.R:1-7
// by factoring a matrix that has extra rows under the training block: the y row turns into (L^-1 y)', rows
// holding K(new, train) turn into (L^-1 K.tp)', and -- when the matrix also has the K(new,new) columns --
// the trailing update leaves the Schur complement there, which the same loop then factors.
#include "common.cuh"
#include <atomic>
#include "ctx.h"
#include <functional>

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Two-level right-looking schedule.  Column tiles (128 wide) are grouped into outer panels of `outer` tiles:
//   inside a panel   potf2(k) -> panel solve(k) -> update of the panel's remaining columns (K = 128)
//   after the panel  one trailing update of everything to the right with K = outer*128
// The wide K is what keeps the DMMA pipe busy: every 128x128 C tile is read and written once per 2*128*128*K
// flops, and with one CTA per SM that traffic is not overlapped with the MMAs (measured: K = 128 -> 53 % of the
// DMMA peak in the update kernel).  With look-ahead the next panel's columns are updated first on the (high
// priority) panel stream so that its factorisation overlaps the rest of the trailing update on the bulk stream.
int trap_factor(fgp_ctx* ctx, int devslot, const TrapDesc& t, cudaStream_t sPanel, cudaStream_t sBulk, bool lookahead) {
    const int NBk = FGP_TILE;
    const int ncT = cdiv(t.ncols, NBk);                 // column tiles (index space shared with rows)
    const int nrT = cdiv(t.nrows, NBk);
    const int nAT = cdiv(t.holeLo, NBk);                // tiles covering the training block (+ y row)
    const int eT0 = t.holeHi / NBk;                     // first extra-row tile
    const bool hasExtra = t.nrows > t.holeHi;
    const bool colHole = t.colHoleHi > t.colHoleLo || (t.colHoleLo < t.ncols);
    const int colSegTile = (t.colHoleLo < t.ncols) ? t.colHoleHi / NBk : ncT;   // first tile of the second column segment
    FgpTimers* tm = ctx->profile ? &ctx->dev[devslot].timers : nullptr;
    const bool la_pre = lookahead && sBulk != nullptr && sBulk != sPanel && t.batch == 1 && !ctx->profile;
    // Look-ahead path: the panel chain of one outer panel must hide under the trailing update of the previous one.
    // Round-1 measurements with one fixed width per factorisation (tools/single_probe*.py): n = 8192 best with 2 tiles per
    // panel (10.1 ms; 4: 10.2), n = 16384 with 8 (53.9 ms; 2: 58.8), n = 32768 with 16 (372 ms = 31.5 TFLOP/s; 2: 432)
    // Look-ahead schedule of a single factorisation, option "outer_single" = 0: the panel width follows the shape of the
    // work.  A column tile costs the chain ~70 us (diagonal block, panel solve, narrow update) whatever the width, while
    // the trailing update of r remaining tiles does ~0.06 r^2 us of work per 128 columns of K on 148 SMs: it hides the
    // chain only while r > ~35.  So: start narrow (2 tiles: the device is busy after 150 us), widen to 8 (16 for very
    // large matrices) while the update dominates -- wide K is where the DMMA kernel is efficient -- and return to 2-tile
    // panels for the chain-bound tail.  A fixed width can still be forced with the option.
    auto width_at = [&](int k) {
        if (!la_pre) return std::max(1, ctx->outer);
        if (ctx->outer_single > 0) return ctx->outer_single;
        const int r = ncT - k;
        if (ncT < 40) return 2;
        if (k == 0) return 2;
        if (k < 6) return 4;
        if (r > 136) return 16;
        if (r > ctx->la_r8) return 8;
        if (r > ctx->la_r4) return 4;
        return 2;
    };
    (void)colHole;
    const int pdlMode = (ctx->pdl && !ctx->graphs) ? (la_pre ? 2 : 1) : 0;      // not inside a stream capture

    auto col_nb = [&](int k) {
        const int k0 = k * NBk;
        if (k0 < t.colHoleLo) return std::min(NBk, t.colHoleLo - k0);
        return std::min(NBk, t.ncols - k0);
    };

    GemmArgs base{};
    base.lda = t.ld; base.ldb = t.ld; base.ldc = t.ld;
    base.strideA = t.strideM; base.strideB = t.strideM; base.strideC = t.strideM;
    base.rowMin = 0; base.rowLo = t.holeLo; base.rowHoleHi = t.holeHi; base.nrows = t.nrows;
    base.colMin = 0; base.colLo = t.colHoleLo; base.colHoleHi = t.colHoleHi; base.ncols = t.ncols;
    base.batch = t.batch;
    base.maxTilesPerCta = la_pre ? 1 : 0;
    base.pdl = pdlMode;
    base.tma = ctx->gemm_tma;
    base.flat = ctx->gemm_flat;

    cudaEvent_t evPanel = nullptr, evBulk = nullptr;
    const bool la = lookahead && sBulk != nullptr && sBulk != sPanel && t.batch == 1 && !ctx->profile;
    if (la) {
        evPanel = ctx->dev[devslot].evA;
        evBulk = ctx->dev[devslot].evB;
        // the bulk stream must see everything queued so far on the panel stream (assembly)
        cudaEventRecord(evPanel, sPanel);
        cudaStreamWaitEvent(sBulk, evPanel, 0);
    }
    bool bulkPending = false;

    // participating row tiles of step k:  S = [lo1,hi1) u [lo2,hi2)
    struct RowSet { int lo1, hi1, lo2, hi2; };
    auto rows_of = [&](int k) {
        const bool fresh = k >= t.cpre;
        RowSet r{0, 0, 0, 0};
        if (t.upperRows) {
            r.lo1 = fresh ? k + 1 : nAT; r.hi1 = nAT;
            r.lo2 = eT0; r.hi2 = std::min(eT0 + k + 1, nrT);
        } else if (!fresh) {
            r.lo2 = eT0; r.hi2 = nrT;
        } else {
            r.lo1 = k + 1; r.hi1 = nrT;
        }
        if (r.hi1 < r.lo1) r.hi1 = r.lo1;
        if (r.hi2 < r.lo2) r.hi2 = r.lo2;
        if (!hasExtra && (t.upperRows || !fresh)) { r.lo2 = r.hi2 = 0; }
        return r;
    };
    // C(I,J) -= X_I X_J'  for J in [cLo,cHi), I in S, I >= J;  X = columns [kA*128, kA*128+K)
    auto make_update = [&](int kA, int K, const RowSet& S, int cLo, int cHi) {
        GemmArgs g = base;
        g.A = t.M + (long)kA * NBk * t.ld;
        g.B = g.A;
        g.C = t.M;
        g.K = K; g.accumulate = 1; g.bTileBase = 0;
        g.tri0 = 0; g.triT = 0; g.rectR0 = 0; g.rectNR = 0; g.rectR1 = 0; g.rectNR1 = 0;
        g.rectC0 = cLo; g.rectNC = cHi - cLo;
        const int r1lo = std::max(S.lo1, cLo), r1hi = S.hi1;
        if (r1hi > r1lo) {
            const int triHi = std::min(r1hi, cHi);
            if (r1lo == cLo) {
                g.tri0 = cLo; g.triT = std::max(0, triHi - cLo);
                g.rectR0 = triHi; g.rectNR = std::max(0, r1hi - triHi);
            } else {
                g.rectR0 = std::max(r1lo, cHi); g.rectNR = std::max(0, r1hi - g.rectR0);
            }
        }
        if (S.hi2 > S.lo2) {
            if (S.lo2 >= cHi) { g.rectR1 = S.lo2; g.rectNR1 = S.hi2 - S.lo2; }
            else if (cLo >= S.lo2) {
                // simulate, prefactored steps, column range inside the new-point block (the look-ahead split hands the
                // panel stream [cLo, cHi) with cHi < ncT): its triangle plus the extra rows below it
                g.tri0 = cLo; g.triT = cHi - cLo;
                g.rectC0 = cLo; g.rectNC = cHi - cLo;
                g.rectR1 = cHi; g.rectNR1 = std::max(0, S.hi2 - cHi);
            } else {
                // simulate, prefactored steps: extra rows x (train columns: rectangle) + (new columns: triangle up to the
                // last row -- only valid when the range runs to the end, which the callers guarantee: panels and
                // look-ahead ranges never straddle the boundary between the two column segments)
                g.rectNC = std::max(0, std::min(cHi, S.lo2) - cLo);
                g.rectR1 = S.lo2; g.rectNR1 = S.hi2 - S.lo2;
                if (cHi > S.lo2) { g.tri0 = std::max(S.lo2, cLo); g.triT = cHi - g.tri0; }
            }
        }
        return g;
    };
    // algorithmic flops of an update launch: the triangle region counts its entries on and below the diagonal only
    // (the kernel skips the warps above it), rectangles in full
    auto flops_of = [&](const GemmArgs& g) {
        const double triRows = (double)g.triT * NBk;
        const double entries = triRows * (triRows + 1.0) / 2.0 + (double)(g.rectNR + g.rectNR1) * g.rectNC * NBk * NBk;
        return entries * 2.0 * (double)g.K * t.batch;
    };

    // Prefactored steps whose column range straddles the boundary between kept and new columns AND whose extra rows reach
    // below the last column tile (fgp_premats_reuse with n a multiple of 128: the y row sits alone in tile ncT): the
    // rectangle (rows below the triangle) x (new columns) is not part of make_update's one-rectangle region -- second launch
    auto make_update_tail = [&](int kA, int K, const RowSet& S, int cLo, int cHi, GemmArgs& g) {
        if (!(S.hi2 > S.lo2) || S.lo2 >= cHi || cLo >= S.lo2 || S.hi2 <= cHi) return false;
        g = base;
        g.A = t.M + (long)kA * NBk * t.ld;
        g.B = g.A;
        g.C = t.M;
        g.K = K; g.accumulate = 1; g.bTileBase = 0;
        g.tri0 = 0; g.triT = 0; g.rectR0 = 0; g.rectNR = 0;
        g.rectR1 = cHi; g.rectNR1 = S.hi2 - cHi;
        g.rectC0 = S.lo2; g.rectNC = cHi - S.lo2;
        return true;
    };
    auto launch_update = [&](int kA, int K, const RowSet& S, int cLo, int cHi, cudaStream_t st, bool timed) {
        GemmArgs g = make_update(kA, K, S, cLo, cHi);
        if (timed && tm) tm->begin(st);
        launch_gemm(g, st);
        if (timed && tm) tm->end(st, FgpTimers::UPDATE, flops_of(g));
        GemmArgs g2;
        if (make_update_tail(kA, K, S, cLo, cHi, g2)) {
            if (timed && tm) tm->begin(st);
            launch_gemm(g2, st);
            if (timed && tm) tm->end(st, FgpTimers::UPDATE, flops_of(g2));
        }
    };
    int k = 0;
    while (k < ncT) {
        // ---- outer panel [k, kend): never straddles the prefactored boundary or the column hole
        const int outer = width_at(k);
        int kend = std::min(k + outer, ncT);
        if (k < t.cpre) kend = std::min(kend, t.cpre);
        if (k < colSegTile) kend = std::min(kend, colSegTile);
        int Kpanel = 0;
        for (int kk = k; kk < kend; ++kk) Kpanel += std::max(0, col_nb(kk));
        // recursive panel: a range of more than one tile is split in halves; the right half is updated with the whole
        // left half at once (K = 128 * tiles), so narrow-K updates only ever touch narrow column ranges
        std::function<void(int, int)> factor_range = [&](int a, int b) {
            if (b - a > 1) {
                const int mid = a + (b - a + 1) / 2;
                factor_range(a, mid);
                int Kl = 0;
                for (int kk = a; kk < mid; ++kk) Kl += std::max(0, col_nb(kk));
                if (Kl > 0) launch_update(a, Kl, rows_of(mid - 1), mid, b, sPanel, true);
                factor_range(mid, b);
                return;
            }
            const int kk = a;
            const int k0 = kk * NBk;
            const int nb = col_nb(kk);
            if (nb <= 0) return;
            const bool fresh = kk >= t.cpre;
            // -- diagonal block
            if (fresh) {
                const int infoOff = (k0 >= t.colHoleHi && t.colHoleLo < t.ncols) ? (k0 - t.colHoleHi) : k0;
                if (tm) tm->begin(sPanel);
                launch_potf2(t.M, t.ld, t.strideM, t.batch, k0, nb, t.W, t.strideW, t.info, infoOff, sPanel, pdlMode);
                if (tm) tm->end(sPanel, FgpTimers::PANEL, 0.0);
            }
            const RowSet S = rows_of(kk);
            // -- panel solve X = P * W_k'  (tile kk itself takes part when its block is ragged: rows below k0+nb)
            GemmArgs g = base;
            g.A = t.M + (long)k0 * t.ld;
            g.B = t.W + (long)kk * NBk * NBk; g.ldb = NBk; g.strideB = t.strideW;
            g.C = t.M + (long)k0 * t.ld;
            g.K = nb; g.accumulate = 0; g.bLowerTri = 1;
            g.tri0 = 0; g.triT = 0;
            int tlo1 = S.lo1, thi1 = S.hi1;
            if (fresh && nb < NBk && (long)k0 + nb < (long)std::min(t.nrows, (kk + 1) * NBk)) {
                if (thi1 > tlo1) tlo1 = kk; else { tlo1 = kk; thi1 = kk + 1; }
            }
            g.rectR0 = tlo1; g.rectNR = thi1 - tlo1; g.rectR1 = S.lo2; g.rectNR1 = S.hi2 - S.lo2;
            g.rectC0 = kk; g.rectNC = 1; g.bTileBase = kk;
            g.rowMin = k0 + nb;
            g.colMin = 0; g.colLo = nb; g.colHoleHi = nb; g.ncols = nb;
            if (g.rectNR + g.rectNR1 > 0) {
                if (tm) tm->begin(sPanel);
                launch_gemm(g, sPanel);
                if (tm) tm->end(sPanel, FgpTimers::TRSM, 0.0);
            }
        };
        factor_range(k, kend);
        // ---- trailing update with the whole panel
        if (kend < ncT && Kpanel > 0) {
            const RowSet S = rows_of(kend - 1);
            if (!la) {
                launch_update(k, Kpanel, S, kend, ncT, sPanel, true);
            } else {
                // look-ahead: the next panel's columns on the panel stream (critical path), the rest on the bulk stream
                int kend2 = std::min(kend + width_at(kend), ncT);
                if (kend < t.cpre) kend2 = std::min(kend2, t.cpre);
                if (kend < colSegTile) kend2 = std::min(kend2, colSegTile);
                cudaEventRecord(evPanel, sPanel);                           // panel solved
                if (bulkPending) cudaStreamWaitEvent(sPanel, evBulk, 0);   // its columns carry the previous bulk update
                launch_update(k, Kpanel, S, kend, kend2, sPanel, false);
                if (kend2 < ncT) {
                    cudaStreamWaitEvent(sBulk, evPanel, 0);
                    launch_update(k, Kpanel, S, kend2, ncT, sBulk, false);
                    cudaEventRecord(evBulk, sBulk);
                    bulkPending = true;
                }
            }
        }
        k = kend;
    }
    if (la && bulkPending) cudaStreamWaitEvent(sPanel, evBulk, 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fgp_set_err_cuda(ctx, "trap_factor launch", e, __FILE__, __LINE__);
        return FGP_ERR_CUDA_;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// nll / sigma^2 from the factored trapezoid (R/3_training_SF.R:250-255, 266-269)
namespace {

__device__ __forceinline__ double block_sum(double v, double* red) {
    // deterministic: fixed tree
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
        red[32] = s;
    }
    __syncthreads();
    return red[32];
}

__global__ void nll_reduce_kernel(const double* M, long ld, long strideM, int n, int yrow, const double* var_known,
                                  const int* info, double* nll, double* sig2) {
    __shared__ double red[33];
    const double* Mb = M + (long)blockIdx.x * strideM;
    double slog = 0.0, sq = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        slog += log(Mb[(long)i + (long)i * ld]);
        const double z = Mb[(long)yrow + (long)i * ld];
        sq = fma(z, z, sq);
    }
    const double sumlog = block_sum(slog, red);
    const double quad = block_sum(sq, red);
    if (threadIdx.x == 0) {
        const double s2 = var_known ? *var_known : quad / (double)n;
        // -llik = 0.5 * (n log(2 pi sig2) + 2 sum(log diag U) + n)      (quirk Q1: '+ n' also with a fixed variance)
        double v = 0.5 * ((double)n * log(2.0 * 3.14159265358979323846 * s2) + 2.0 * sumlog + (double)n);
        if (info && info[blockIdx.x] != 0) v = nan("");
        nll[blockIdx.x] = v;
        sig2[blockIdx.x] = s2;
    }
}

// Row statistics of the solved extra rows, one pass over the data:
//   norm[b*nrow + i] = sum_j M_b(row0+i, j)^2            (sd of predict, diag(R^-1) of LOOCV)
//   dot [b*nrow + i] = sum_j M_b(row0+i, j) z_b[j*zstride]   (predictive mean, R^-1 y)
// Block = 32 consecutive rows (lanes: coalesced along the column-major rows) x 32 warps that interleave the columns;
// the 32 partial sums of a row are added in a fixed order, so results do not depend on the launch shape.
// jstart_by_row (LOOCV, rows of L^-T): columns left of the row's own 128-tile are known zeros and skipped.
__global__ void __launch_bounds__(1024) rowstats_kernel(const double* M, long ld, long strideM, int row0, int nrow, int ncol,
                                                        int jstart_by_row, const double* z, long zstride, long zbatch,
                                                        double* norm, double* dot) {
    __shared__ double redN[32][33], redD[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const double* Mb = M + (long)blockIdx.y * strideM;
    const double* zb = z ? z + (long)blockIdx.y * zbatch : nullptr;
    const int j0 = jstart_by_row ? ((blockIdx.x * 32) / FGP_TILE) * FGP_TILE : 0;
    double sn = 0.0, sd = 0.0;
    if (i < nrow) {
        const double* rowp = Mb + (long)(row0 + i);
        for (int j = j0 + w; j < ncol; j += 32) {
            const double v = rowp[(long)j * ld];
            sn = fma(v, v, sn);
            if (zb) sd = fma(v, zb[(long)j * zstride], sd);
        }
    }
    redN[w][lane] = sn; redD[w][lane] = sd;
    __syncthreads();
    if (w == 0 && i < nrow) {
        double a = 0.0, c = 0.0;
        for (int q = 0; q < 32; ++q) { a += redN[q][lane]; c += redD[q][lane]; }
        if (norm) norm[(long)blockIdx.y * nrow + i] = a;
        if (dot) dot[(long)blockIdx.y * nrow + i] = c;
    }
}

// sims(s, i) = mean[i] + sum_{j<=i} Lc(i,j) Z(j,s)    (R/5_simulation_SF.R:14-15), sims is nsim x m col-major
__global__ void lower_times_kernel(const double* Lc, long ld, int m, const double* Z, int nsim, const double* mean,
                                   double* sims) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int s = blockIdx.y; s < nsim; s += gridDim.y) {
        double acc = 0.0;
        const double* z = Z + (long)s * m;
        for (int j = 0; j <= i; ++j) acc = fma(Lc[(long)i + (long)j * ld], z[j], acc);
        sims[(long)s + (long)i * nsim] = mean[i] + acc;
    }
}

__global__ void extract_lower_kernel(const double* M, long ld, int n, double* out) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n * n) return;
    const int r = (int)(idx % n), c = (int)(idx / n);
    out[idx] = (r >= c) ? M[(long)r + (long)c * ld] : 0.0;
}

__global__ void fill_yrow_kernel(double* M, long ld, long strideM, int row, const double* y, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) M[(long)blockIdx.y * strideM + (long)row + (long)j * ld] = y[j];
}
// heterogeneous batch: every element brings its own output vector
__global__ void fill_yrow_elems_kernel(double* M, long ld, long strideM, int row, const AsmElem* elems, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) M[(long)blockIdx.y * strideM + (long)row + (long)j * ld] = elems[blockIdx.y].y[j];
}

// identity rows for LOOCV: rows row0 .. row0+n-1 (all columns: the wide trailing update reads rows that have not
// started yet and needs zeros there)
__global__ void fill_identity_rows_kernel(double* M, long ld, long strideM, int row0, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // extra row index
    if (i >= n) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y)         // column
        M[(long)blockIdx.z * strideM + (long)(row0 + i) + (long)j * ld] = (i == j) ? 1.0 : 0.0;
}


// zero the strict upper triangle of an m x m block (the factor of the conditional covariance becomes a dense operand)
__global__ void zero_strict_upper_kernel(double* C, long ld, int m) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    for (int c = blockIdx.y + r + 1; c < m; c += gridDim.y) C[(long)r + (long)c * ld] = 0.0;
}

// out(s, i) = v[i] for all s < nrow   (sims start as the predictive mean, R/5_simulation_SF.R:15)
__global__ void bcast_rows_kernel(double* out, long ld, int nrow, int ncol, const double* v) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)nrow * ncol) return;
    const int s = (int)(idx % nrow), i = (int)(idx / nrow);
    out[(long)s + (long)i * ld] = v[i];
}

// sigma^2-only update of a stored model: L <- a L (lower trapezoid), (L^-1 y)' <- (L^-1 y)' / a, inv(L_kk) <- inv(L_kk) / a
__global__ void scale_factor_kernel(double* M, long ld, int n, double a, double ainv) {
    const int j = blockIdx.y;                                  // column
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += gridDim.x * blockDim.x) {
        if (r < j) continue;
        double* e = M + (long)r + (long)j * ld;
        *e = (r == n) ? *e * ainv : *e * a;
    }
}
__global__ void scale_vec_kernel(double* v, long count, double a) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) v[i] *= a;
}

__global__ void copy_row_kernel(double* M, long ld, int ncol, int src, int dst) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ncol) M[(long)dst + (long)j * ld] = M[(long)src + (long)j * ld];
}

// W_k = inv(L_kk) for every 128 x 128 diagonal block of a stored factor (restore of a saved model: the blocks the
// panel solves of predict / simulate multiply with).  One CTA per block, thread c = column c by forward substitution;
// a ragged last block is padded with the identity, as potf2_inv_kernel does.
__global__ void __launch_bounds__(FGP_TILE) trtri_blocks_kernel(const double* M, long ld, int n, double* W) {
    extern __shared__ double Ls[];         // [i][j], pitch 129
    const int k0 = blockIdx.x * FGP_TILE, nb = min(FGP_TILE, n - k0), c = threadIdx.x;
    for (int j = 0; j < FGP_TILE; ++j) {   // coalesced along rows
        double v = (c == j) ? 1.0 : 0.0;
        if (c < nb && j < nb && c >= j) v = M[(long)(k0 + c) + (long)(k0 + j) * ld];
        Ls[c * 129 + j] = v;
    }
    __syncthreads();
    // column c of inv(L): x_i = (delta_ic - sum_{j<i} L_ij x_j) / L_ii.  Every thread walks the same (i, j) range (its
    // leading entries are zeros), so the L reads are broadcasts and the private x vector sits in coalesced local memory
    double x[FGP_TILE];
    for (int i = 0; i < FGP_TILE; ++i) {
        double s = (i == c) ? 1.0 : 0.0;
        for (int j = 0; j < i; ++j) s = fma(-Ls[i * 129 + j], x[j], s);
        x[i] = (i < c) ? 0.0 : s / Ls[i * 129 + i];
    }
    double* Wb = W + (long)blockIdx.x * FGP_TILE * FGP_TILE;
    __syncthreads();
    for (int i = 0; i < FGP_TILE; ++i) Ls[c * 129 + i] = x[i];                               // Ls[c][i] = X(i, c)
    __syncthreads();
    for (int col = 0; col < FGP_TILE; ++col) Wb[c + col * FGP_TILE] = Ls[col * 129 + c];     // W(r = c, col), coalesced
}

}  // namespace

void launch_nll_reduce(const double* M, long ld, long strideM, int batch, int n, int yrow, const double* var_known,
                       const int* info, double* nll, double* sig2, cudaStream_t st) {
    FGP_COUNT_LAUNCH();
    nll_reduce_kernel<<<batch, 256, 0, st>>>(M, ld, strideM, n, yrow, var_known, info, nll, sig2);
}
void launch_rownorm2(const double* M, long ld, long strideM, int batch, int row0, int nrow, int ncol, int upper, double* out,
                     cudaStream_t st) {
    if (nrow <= 0 || batch <= 0) return;
    FGP_COUNT_LAUNCH();
    rowstats_kernel<<<dim3(cdiv(nrow, 32), batch), 1024, 0, st>>>(M, ld, strideM, row0, nrow, ncol, upper, nullptr, 0, 0, out, nullptr);
}
void launch_rowgemv(const double* M, long ld, long strideM, int batch, int row0, int nrow, int ncol, int upper,
                    const double* z, long zstride, long zbatch, double* out, cudaStream_t st) {
    if (nrow <= 0 || batch <= 0) return;
    FGP_COUNT_LAUNCH();
    rowstats_kernel<<<dim3(cdiv(nrow, 32), batch), 1024, 0, st>>>(M, ld, strideM, row0, nrow, ncol, upper, z, zstride, zbatch, nullptr, out);
}
void launch_rowstats(const double* M, long ld, long strideM, int batch, int row0, int nrow, int ncol, int upper,
                     const double* z, long zstride, long zbatch, double* norm, double* dot, cudaStream_t st) {
    if (nrow <= 0 || batch <= 0) return;
    FGP_COUNT_LAUNCH();
    rowstats_kernel<<<dim3(cdiv(nrow, 32), batch), 1024, 0, st>>>(M, ld, strideM, row0, nrow, ncol, upper, z, zstride, zbatch, norm, dot);
}
void launch_lower_times(const double* Lc, long ld, int m, const double* Z, int nsim, const double* mean, double* sims,
                        cudaStream_t st) {
    dim3 grid(cdiv(m, 128), std::min(nsim, 65535));
    FGP_COUNT_LAUNCH();
    lower_times_kernel<<<grid, 128, 0, st>>>(Lc, ld, m, Z, nsim, mean, sims);
}
void launch_extract_lower(const double* M, long ld, int n, double* out, cudaStream_t st) {
    const long tot = (long)n * n;
    FGP_COUNT_LAUNCH();
    extract_lower_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(M, ld, n, out);
}
void launch_fill_yrow(double* M, long ld, long strideM, int batch, int row, const double* y, int n, cudaStream_t st) {
    dim3 grid(cdiv(n, 256), batch);
    FGP_COUNT_LAUNCH();
    fill_yrow_kernel<<<grid, 256, 0, st>>>(M, ld, strideM, row, y, n);
}
void launch_fill_yrow_elems(double* M, long ld, long strideM, int batch, int row, const AsmElem* elems, int n, cudaStream_t st) {
    dim3 grid(cdiv(n, 256), batch);
    FGP_COUNT_LAUNCH();
    fill_yrow_elems_kernel<<<grid, 256, 0, st>>>(M, ld, strideM, row, elems, n);
}
void launch_fill_identity_rows(double* M, long ld, long strideM, int batch, int row0, int n, cudaStream_t st) {
    dim3 grid(cdiv(n, 128), std::min(n, 65535), batch);
    FGP_COUNT_LAUNCH();
    fill_identity_rows_kernel<<<grid, 128, 0, st>>>(M, ld, strideM, row0, n);
}

void launch_zero_strict_upper(double* C, long ld, int m, cudaStream_t st) {
    if (m <= 1) return;
    FGP_COUNT_LAUNCH();
    zero_strict_upper_kernel<<<dim3(cdiv(m, 128), std::min(m, 64)), 128, 0, st>>>(C, ld, m);
}
void launch_bcast_rows(double* out, long ld, int nrow, int ncol, const double* v, cudaStream_t st) {
    const long tot = (long)nrow * ncol;
    if (tot <= 0) return;
    FGP_COUNT_LAUNCH();
    bcast_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(out, ld, nrow, ncol, v);
}
void launch_scale_factor(double* M, long ld, int n, double* W, long wcount, double a, cudaStream_t st) {
    FGP_COUNT_LAUNCH();
    scale_factor_kernel<<<dim3(std::min(cdiv(n + 1, 256), 64), n), 256, 0, st>>>(M, ld, n, a, 1.0 / a);
    FGP_COUNT_LAUNCH();
    scale_vec_kernel<<<(unsigned)((wcount + 255) / 256), 256, 0, st>>>(W, wcount, 1.0 / a);
}
void launch_copy_row(double* M, long ld, int ncol, int src, int dst, cudaStream_t st) {
    FGP_COUNT_LAUNCH();
    copy_row_kernel<<<cdiv(ncol, 256), 256, 0, st>>>(M, ld, ncol, src, dst);
}
void launch_trtri_blocks(const double* M, long ld, int n, double* W, cudaStream_t st) {
    static std::atomic<bool> attr_set[FGP_MAXDEV];        // zero-initialised; several host threads may arrive at once
    const int smem = FGP_TILE * 129 * (int)sizeof(double);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < FGP_MAXDEV && !attr_set[dev]) {
        cudaFuncSetAttribute(trtri_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set[dev] = true;
    }
    FGP_COUNT_LAUNCH();
    trtri_blocks_kernel<<<cdiv(n, FGP_TILE), FGP_TILE, smem, st>>>(M, ld, n, W);
}

int g_plain_tma = 1;      // operand path of launch_gemm_plain (set with the context option "gemm_tma")

// C (rows x cols of 128-tiles, lower triangle only when `tri`) -= A B' over K columns: plain operands without holes.
// rows / cols are the valid extents (ragged last tiles are masked).
void launch_gemm_plain(const double* A, long lda, const double* B, long ldb, double* C, long ldc, int rows, int cols, int K,
                       bool tri, cudaStream_t st) {
    GemmArgs g{};
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    g.K = K; g.accumulate = 1; g.batch = 1; g.tma = g_plain_tma;
    if (tri) { g.tri0 = 0; g.triT = cdiv(rows, FGP_TILE); }
    else { g.rectR0 = 0; g.rectNR = cdiv(rows, FGP_TILE); g.rectC0 = 0; g.rectNC = cdiv(cols, FGP_TILE); }
    g.rowMin = 0; g.rowLo = rows; g.rowHoleHi = rows; g.nrows = rows;
    g.colMin = 0; g.colLo = cols; g.colHoleHi = cols; g.ncols = cols;
    launch_gemm(g, st);
}
