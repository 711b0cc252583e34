// Internal declarations shared by the translation units of libfungp_b200.so (sm_100a only).
// Layout of one factorisation workspace ("trapezoid"), column-major doubles, leading dimension ld:
//
//        cols 0 .. ncols-1
//   rows 0 .. nA-1          training block (lower triangle used; tiles of 128)
//   row  nA (optional)      the y row: after the factorisation it holds (L^-1 y)'      [fit / premats / loocv]
//   rows holeLo .. holeHi-1 padding up to the next multiple of 128 (never read or written)
//   rows holeHi .. nrows-1  extra rows: cross-covariance K(new, train) -> (L^-1 K)' ; for simulate the matrix
//                           also has columns holeHi.. for K(new,new) (Schur complement, then its own Cholesky);
//                           for LOOCV identity rows -> L^-T.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <mutex>

#define FGP_TILE 128            // factorisation tile (panel width NB and GEMM CTA tile)
#define FGP_MAXDEV 16

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            fgp_set_err_cuda(ctx_for_err, #expr, _e, __FILE__, __LINE__);                       \
            return FGP_ERR_CUDA_;                                                               \
        }                                                                                       \
    } while (0)

enum { FGP_ERR_ARG_ = -1, FGP_ERR_CUDA_ = -2, FGP_ERR_NODEVICE_ = -3, FGP_ERR_STATE_ = -4, FGP_ERR_ALLOC_ = -5 };

struct fgp_ctx;
void fgp_set_err(fgp_ctx* ctx, const std::string& msg);
void fgp_set_err_cuda(fgp_ctx* ctx, const char* what, cudaError_t e, const char* file, int line);

// ---------------------------------------------------------------------------------------------------------
// coordinate "units" consumed by the assembly kernels (one per squared-difference term)
//   kind 0: index unit  -- one column x;            d = |x_i - x_j|            (L2_byindex, scalar inputs)
//   kind 1: group unit  -- one column z_a of the whitened coefficients z = c W, J = W W';  d^2 += (z_ia-z_ja)^2  (L2_bygroup)
// `last` marks the unit that closes its length-scale term.
struct FgpUnit {
    int kind;
    int col;   // first column in the coordinate matrix
    int term;  // length-scale index
    int last;
};

// one element of a heterogeneous batch (fitlock.cu: the pending likelihood points of MANY candidate structures in one
// launch): its own coordinate matrix, unit table, length-scale multipliers and output vector.  All elements share n.
struct AsmElem {
    const double* X;          // n x ncoord, leading dimension = AsmArgs::ldP
    const FgpUnit* units;
    const double* scale;      // nterms multipliers of THIS element
    const double* y;          // output vector for the y row
    int nunits;
    int pad_;
};

struct AsmArgs {
    // row points P, column points Q (column-major coordinate matrices, leading dims ldP / ldQ)
    const double* XP; int nP; long ldP;
    const double* XQ; int nQ; long ldQ;
    const FgpUnit* units; int nunits;
    const double* scale;      // per batch element: nterms multipliers c/theta_t  (c = 1, sqrt5, sqrt3)
    int nterms;
    int ker;                  // 1 gauss 2 matern5_2 3 matern3_2
    int sym;                  // 1: P == Q, only tiles with i-tile >= j-tile are produced
    double mult;              // entries are mult * (r + [i==j] * diag_rel) + [i==j] * diag_abs
    double diag_rel, diag_abs;
    double* M; long ld; long strideM;   // destination (+ batch stride); element (rowOff+i, colOff+j)
    int rowOff, colOff;
    int mirror;               // sym only: also write the transposed tile (full symmetric output)
    const AsmElem* elems;     // non-null: batch element b takes XP = XQ, units and scale from elems[b] (sym only)
};

void launch_assemble(const AsmArgs& a, int batch, cudaStream_t st);
// batch-fused assembly of symmetric training matrices: the elements of a group share structure and coordinates and
// differ in their length-scales only (assemble.cu)
struct AsmGroup { int first, count; };
bool assemble_fusable(int nterms, int nunits);
void launch_assemble_fused(const AsmArgs& a, const AsmGroup* groups, int ngroups, int count, cudaStream_t st);
// max over pairs of the group distance of one L2_bygroup term / materialise one distance matrix
void launch_distmax(const double* X, int n, long ldX, const FgpUnit* units, int u0, int u1, double* out_max,
                    cudaStream_t st);
void launch_distmat(const double* X, int n, long ldX, const FgpUnit* units, int u0, int u1, double* out, long ld,
                    cudaStream_t st);
void launch_fill_yrow(double* M, long ld, long strideM, int batch, int row, const double* y, int n, cudaStream_t st);
void launch_fill_yrow_elems(double* M, long ld, long strideM, int batch, int row, const AsmElem* elems, int n, cudaStream_t st);
void launch_fill_identity_rows(double* M, long ld, long strideM, int batch, int row0, int n, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------
// factorisation engine
struct TrapDesc {
    double* M; long ld; long strideM; int batch;
    int ncols;      // columns to factor through (multiple tiles of 128; last may be ragged)
    int nrows;      // total rows incl. extra rows
    int holeLo;     // rows [holeLo, holeHi) do not exist
    int holeHi;
    int colHoleLo;  // same for columns (simulate: train cols [0,n), new cols [holeHi, ...)); = ncols when none
    int colHoleHi;
    int cpre;       // number of leading column tiles that are already factored (predict / simulate)
    int upperRows;  // 1: the extra rows start as identity and only extra-row tiles I' <= k are touched (LOOCV)
    double* W; long strideW;  // inverse diagonal blocks, 128 x 128 each, tile k at W + k*128*128 (+ b*strideW)
    int* info;      // per batch element, 0 or failing minor order
};

struct FgpTimers;
int trap_factor(fgp_ctx* ctx, int devslot, const TrapDesc& t, cudaStream_t sPanel, cudaStream_t sBulk, bool lookahead);

void launch_potf2(double* M, long ld, long strideM, int batch, int k0, int nb, double* W, long strideW, int* info,
                  int infoBase, cudaStream_t st, int pdl = 0);

// the whole factorisation of a batch as one persistent dataflow kernel (chol_dag.cu): any n; nrows == n + 1 (the y row
// rides along) or nrows == n (hasY = 0); and its solve mode for the extra rows of predict / simulate
bool chol_dag_supported(int n, int nrows);
size_t chol_dag_flag_bytes(int n, int batch);
size_t chol_dag_solve_flag_bytes(int n, int mExtra);
bool launch_chol_dag(double* M, long ld, long strideM, int batch, int n, double* W, long strideW, int* info, void* flags,
                     cudaStream_t st, int group = 1, int matrixMajor = 1, int hasY = 1);
bool launch_chol_dag_solve(double* M, long ld, int n, int rowOff, int mExtra, double* W, void* flags, cudaStream_t st, int upper = 0);

struct GemmArgs {
    const double* A; long lda; long strideA;   // row operand: element (r, kk) = A[r + kk*lda], r = global row
    const double* B; long ldb; long strideB;   // col operand: element (c, kk) = B[c + kk*ldb]
    double* C; long ldc; long strideC;         // C(r, c) = C[r + c*ldc]
    int K;                                     // valid k extent
    int accumulate;                            // 1: C -= A B'   0: C = A B'
    // tile region: triangle over tiles [tri0, tri0+triT) (J <= I) followed by rectangle
    int tri0, triT;
    // rectangle rows: tiles [rectR0, rectR0+rectNR) then [rectR1, rectR1+rectNR1); cols [rectC0, rectC0+rectNC)
    int rectR0, rectNR, rectR1, rectNR1, rectC0, rectNC;
    int bTileBase;                             // B rows / C cols of tile J start at (J - bTileBase)*128
    // validity of global rows r and of B-row / C-col indices c
    int rowMin, rowLo, rowHoleHi, nrows;       // valid r: r >= rowMin && (r < rowLo || (rowHoleHi <= r < nrows))
    int colMin, colLo, colHoleHi, ncols;       // same for c
    int batch;
    int bLowerTri;                             // 1: B(c, k) = 0 for k > c (panel solve with inv(L_kk)): skip the zero half
    int maxTilesPerCta;                        // 0 = automatic; 1 = one tile per CTA (look-ahead: the panel stream must find a free SM quickly)
    int dbg;                                   // experiment switches, honoured only when built with -DFGP_GEMM_DEBUG
    int tma;                                   // 1: operands by TMA (gemm_tma.cu), 0: per-thread cp.async (gemm_dmma.cu)
    int flat;                                  // TMA kernel: > 0 = walk all (tile, element) pairs with one CTA per SM when there are at least
                                               // flat x SMs of them and K <= 1024 (batches of small matrices); 0 = never
    int pdl;                                   // 1: programmatic dependent launch (the prologue overlaps the previous kernel's tail); 2: also
                                               // release the dependents at once (single-factorisation chain: the next kernel waits resident)
};
void launch_gemm(const GemmArgs& g, cudaStream_t st);

// logdet / quadratic form reduction -> nll, sig2  (R/3_training_SF.R:250-255)
void launch_nll_reduce(const double* M, long ld, long strideM, int batch, int n, int yrow, const double* var_known,
                       const int* info, double* nll, double* sig2, cudaStream_t st);
// generic small kernels used by predict / simulate / loocv
// out[b*nrow + i] = sum_j M_b(row0+i, j)^2 and sum_j M_b(row0+i, j) z_b[j*zstride]; one warp per row, fixed summation
// order; `upper`: columns before the row's own tile are known zeros (LOOCV's L^-T) and skipped
void launch_rownorm2(const double* M, long ld, long strideM, int batch, int row0, int nrow, int ncol, int upper, double* out,
                     cudaStream_t st);
void launch_rowgemv(const double* M, long ld, long strideM, int batch, int row0, int nrow, int ncol, int upper,
                    const double* z, long zstride, long zbatch, double* out, cudaStream_t st);
void launch_rowstats(const double* M, long ld, long strideM, int batch, int row0, int nrow, int ncol, int upper,
                     const double* z, long zstride, long zbatch, double* norm, double* dot, cudaStream_t st);
void launch_lower_times(const double* Lc, long ld, int m, const double* Z, int nsim, const double* mean, double* sims,
                        cudaStream_t st);
void launch_zero_strict_upper(double* C, long ld, int m, cudaStream_t st);
void launch_bcast_rows(double* out, long ld, int nrow, int ncol, const double* v, cudaStream_t st);
void launch_scale_factor(double* M, long ld, int n, double* W, long wcount, double a, cudaStream_t st);
void launch_copy_row(double* M, long ld, int ncol, int src, int dst, cudaStream_t st);
void launch_trtri_blocks(const double* M, long ld, int n, double* W, cudaStream_t st);
void launch_gemm_plain(const double* A, long lda, const double* B, long ldb, double* C, long ldc, int rows, int cols, int K,
                       bool tri, cudaStream_t st);
void launch_extract_lower(const double* M, long ld, int n, double* out, cudaStream_t st);

int fgp_peaks_device(double* dmma_tflops, double* dfma_tflops, double* copy_gbs);

// number of kernels of this library launched by the process (bench.py reports it as gpu_launches)
#include <atomic>
extern std::atomic<long long> g_fgp_launches;
#define FGP_COUNT_LAUNCH() (g_fgp_launches.fetch_add(1, std::memory_order_relaxed))
// work items of the process (bench.py turns them into aggregate FLOP/s): [0] likelihood factorisations (n^3/3 each),
// [1] leave-one-out factorisations (Cholesky + L^-T: 2 n^3/3), [2] preMats factorisations, [3] candidates scored
extern std::atomic<long long> g_fgp_work[4];
#define FGP_COUNT_WORK(i, k) (g_fgp_work[i].fetch_add((k), std::memory_order_relaxed))
