// Model state behind the C ABI: preMats, predict, simulate, LOOCV, and what comes after a fit -- (de)serialisation of the
// factor, large host formats, and the fast paths of update().  Reference: R/4_prediction_SF.R:2-32,
// R/5_simulation_SF.R:3-23, R/8_outilsStats.R:1-7, R/6_updating.R:7-643, R/7_plottingFunctions.R:9-11.
//
// The factor of sig2 (R + nugget I) stays on the device inside the problem, in a buffer whose leading dimension leaves
// `predict_reserve` free rows under it: predict and simulate assemble K(new, train) straight into those rows and solve
// them against the resident factor IN PLACE (the trapezoid engine of chol.cu with every column tile prefactored), so
// no copy of L is made per call.  More new points than reserved rows are processed in chunks (predictions of different
// points are independent); only a simulate call with more points than rows falls back to a private copy.
#include <sys/mman.h>
#include "ctx.h"
#include "../../include/fungp_b200.h"
#include <cstring>
#include <thread>

#define CK(expr)                                                              \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            fgp_set_err_cuda(ctx, #expr, _e, __FILE__, __LINE__);             \
            return FGP_ERR_CUDA;                                              \
        }                                                                     \
    } while (0)

static inline cudaError_t wait_stream(fgp_ctx* ctx, DevState& d, cudaStream_t st) { return fgp_wait_stream(ctx, d, st); }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long roundup(long a, long b) { return fgp_roundup(a, b); }
static inline long pick_ld(long rows) { return fgp_pick_ld(rows); }

// ------------------------------------------------------------------------------------------------ large host formats
// Column-major matrices between pageable host memory and the device through two pinned staging buffers: while one chunk
// of columns is on the bus, a few host threads move the previous one between the staging buffer and the caller's array
// (a single pageable cudaMemcpy of preMats$L at n = 32768 -- 8.6 GB -- cost more than the factorisation).
// lower = 1: only rows >= column index travel; on the way out the strict upper part of the host matrix is zero-filled
// (R's preMats$L carries explicit zeros), on the way in it is ignored.
namespace {

const size_t STAGE_BYTES = (size_t)48 << 20;

int ensure_stage(fgp_ctx* ctx, DevState& d) {
    if (d.pinnedCap >= 2 * STAGE_BYTES) return 0;
    if (d.pinned) cudaFreeHost(d.pinned);
    d.pinned = nullptr; d.pinnedCap = 0;
    if (cudaHostAlloc(&d.pinned, 2 * STAGE_BYTES, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        fgp_set_err(ctx, "out of pinned host memory (transfer staging)");
        return FGP_ERR_ALLOC;
    }
    d.pinnedCap = 2 * STAGE_BYTES;
    return 0;
}

// run fn(t, nthreads) on a few host threads
template <class F>
void host_parallel(int want, long work_bytes, F fn) {
    unsigned hw = std::thread::hardware_concurrency();
    // writing a fresh destination is bound by page faults per thread: 4 / 8 / 12 / 16 threads move the factor of n = 32768
    // at 11 / 14-17 / 20 / 23 GB/s (tools/lout_probe.py), so all cores up to 16 take part for the length of the copy
    int nt = want > 0 ? want : (int)std::max(1u, std::min(16u, hw));
    if (work_bytes < (4 << 20)) nt = 1;
    if (nt == 1) { fn(0, 1); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back([=]() { fn(t, nt); });
    fn(0, nt);
    for (auto& x : th) x.join();
}

struct Chunk { long c0, c1, r0, rows; };      // columns [c0, c1), rows [r0, r0 + rows)

int copy_matrix(fgp_ctx* ctx, DevState& d, bool toHost, double* host, long ldh, double* dev, long ldd, long rows, long cols,
                int lower, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    int rc = ensure_stage(ctx, d);
    if (rc) return rc;
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    double* stage[2] = {(double*)d.pinned, (double*)((char*)d.pinned + STAGE_BYTES)};
    std::vector<Chunk> chunks;
    for (long c0 = 0; c0 < cols;) {
        const long r0 = lower ? c0 : 0, h = rows - r0;
        if (h <= 0) break;
        const long w = std::max<long>(1, std::min<long>(cols - c0, (long)(STAGE_BYTES / sizeof(double)) / h));
        chunks.push_back({c0, c0 + w, r0, h});
        c0 += w;
    }
    auto host_side = [&](const Chunk& ch, double* buf) {       // staging <-> caller's matrix
        host_parallel(ctx->host_threads, ch.rows * (ch.c1 - ch.c0) * 8, [&](int t, int nt) {
            for (long c = ch.c0 + t; c < ch.c1; c += nt) {
                double* hcol = host + c * ldh;
                double* scol = buf + (c - ch.c0) * ch.rows;
                const long skip = lower ? c - ch.r0 : 0;       // rows of this column above the diagonal inside the chunk
                if (toHost) {
                    if (lower && !ctx->out_prezeroed) memset(hcol, 0, sizeof(double) * (size_t)std::min(c, rows));
                    memcpy(hcol + ch.r0 + skip, scol + skip, sizeof(double) * (size_t)(ch.rows - skip));
                } else {
                    if (skip) memset(scol, 0, sizeof(double) * (size_t)skip);
                    memcpy(scol + skip, hcol + ch.r0 + skip, sizeof(double) * (size_t)(ch.rows - skip));
                }
            }
        });
    };
    auto bus = [&](const Chunk& ch, double* buf) -> cudaError_t {
        double* dptr = dev + ch.r0 + ch.c0 * ldd;
        if (toHost)
            return cudaMemcpy2DAsync(buf, sizeof(double) * ch.rows, dptr, sizeof(double) * ldd, sizeof(double) * ch.rows, ch.c1 - ch.c0,
                                     cudaMemcpyDeviceToHost, st);
        return cudaMemcpy2DAsync(dptr, sizeof(double) * ldd, buf, sizeof(double) * ch.rows, sizeof(double) * ch.rows, ch.c1 - ch.c0,
                                 cudaMemcpyHostToDevice, st);
    };
    const int nc = (int)chunks.size();
    cudaError_t e = cudaSuccess;
    if (toHost && ctx->host_thp && (size_t)cols * ldh * sizeof(double) >= ((size_t)64 << 20)) {
        // The caller's matrix is usually fresh memory (R's allocVector, numpy.empty): writing it is bound by page faults
        // -- 2.1 M of them for the 8.6 GB factor of n = 32768.  Ask for transparent huge pages on the 2 MB-aligned
        // interior; a hint only, ignored where THP is off, harmless on memory that is already populated.
        const uintptr_t lo = (reinterpret_cast<uintptr_t>(host) + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
        const uintptr_t hi = (reinterpret_cast<uintptr_t>(host) + (size_t)cols * ldh * sizeof(double)) & ~(uintptr_t)((2u << 20) - 1);
        if (hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
    }
    if (toHost) {
        for (int i = 0; i < nc + 1 && e == cudaSuccess; ++i) {
            if (i < nc) { e = bus(chunks[i], stage[i & 1]); if (e == cudaSuccess) e = cudaEventRecord(ev[i & 1], st); }
            if (i >= 1 && e == cudaSuccess) {
                e = cudaEventSynchronize(ev[(i - 1) & 1]);
                if (e == cudaSuccess) host_side(chunks[i - 1], stage[(i - 1) & 1]);
            }
        }
        if (lower && !ctx->out_prezeroed && e == cudaSuccess)  // columns that had no rows at all (cols > rows)
            for (long c = rows; c < cols; ++c) memset(host + c * ldh, 0, sizeof(double) * (size_t)rows);
    } else {
        for (int i = 0; i < nc && e == cudaSuccess; ++i) {
            if (i >= 2) e = cudaEventSynchronize(ev[i & 1]);   // the DMA that last read this staging buffer
            if (e != cudaSuccess) break;
            host_side(chunks[i], stage[i & 1]);
            e = bus(chunks[i], stage[i & 1]);
            if (e == cudaSuccess) e = cudaEventRecord(ev[i & 1], st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    for (int i = 0; i < 2; ++i) cudaEventDestroy(ev[i]);
    if (e != cudaSuccess) { fgp_set_err_cuda(ctx, "chunked matrix transfer", e, __FILE__, __LINE__); return FGP_ERR_CUDA; }
    return 0;
}

// factor storage of a problem on device 0: ld leaves the reserve rows free under the (L^-1 y)' row
int ensure_factor_storage(fgp_problem* P) {
    fgp_ctx* ctx = P->ctx;
    DevState& d = ctx->dev[0];
    ProblemDev& pd = P->pd[0];
    if (pd.Lfac) return 0;
    const int n = P->n;
    const long ld = pick_ld(roundup(n + 1, FGP_TILE) + std::max(FGP_TILE, ctx->predict_reserve));
    const int ncT = cdiv(n, FGP_TILE);
    pd.Lfac = (double*)d.pool.get(sizeof(double) * (size_t)ld * n);
    pd.Wall = (double*)d.pool.get(sizeof(double) * (size_t)ncT * FGP_TILE * FGP_TILE);
    if (!pd.Lfac || !pd.Wall) {
        if (pd.Lfac) d.pool.put(pd.Lfac);
        if (pd.Wall) d.pool.put(pd.Wall);
        pd.Lfac = pd.Wall = nullptr;
        fgp_set_err(ctx, "out of device memory (stored factor)");
        return FGP_ERR_ALLOC;
    }
    pd.ldL = ld;
    return 0;
}

void set_model(fgp_problem* P, int ker, double nugget, double sig2, const double* theta) {
    P->hasModel = true; P->ker = ker; P->nugget = nugget; P->sig2 = sig2;
    P->theta.assign(theta, theta + P->nterms);
    P->loocvValid = false;
}

// preMats$L (explicit zeros above the diagonal) and preMats$LInvY to the host
int fetch_premats(fgp_problem* P, double* L_out, double* LInvY_out) {
    fgp_ctx* ctx = P->ctx;
    DevState& d = ctx->dev[0];
    ProblemDev& pd = P->pd[0];
    const int n = P->n;
    cudaStream_t st = d.sPanel[0];
    if (LInvY_out) {
        // the (L^-1 y)' row is strided in the trapezoid: gather it on the device, one contiguous copy out
        if (d.small.ensure(sizeof(double) * (size_t)n)) return FGP_ERR_ALLOC;
        CK(cudaMemcpy2DAsync(d.small.p, sizeof(double), pd.Lfac + n, sizeof(double) * pd.ldL, sizeof(double), n,
                             cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(LInvY_out, d.small.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        CK(wait_stream(ctx, d, st));
    }
    if (L_out) return copy_matrix(ctx, d, true, L_out, n, pd.Lfac, pd.ldL, n, n, 1, st);
    return 0;
}

int upload_new(fgp_problem* P, DevState& d, int m, const double* sNew, const double* const* coefsNew, double** Xnew_dev) {
    fgp_ctx* ctx = P->ctx;
    std::vector<double> X;
    build_coords(P, m, sNew, coefsNew, X);
    if (d.aux.ensure(sizeof(double) * std::max<size_t>(1, X.size()))) return FGP_ERR_ALLOC;
    CK(cudaMemcpyAsync(d.aux.p, X.data(), sizeof(double) * X.size(), cudaMemcpyHostToDevice, d.sPanel[0]));
    CK(wait_stream(ctx, d, d.sPanel[0]));
    *Xnew_dev = (double*)d.aux.p;
    return 0;
}

// scales | info on the device for single-model calls
int stage_scales(fgp_problem* P, DevState& d, int ker, const double* theta, size_t extraDoubles, double** dScale, int** dInfo,
                 double** dExtra, cudaStream_t st) {
    fgp_ctx* ctx = P->ctx;
    const int nt = P->nterms;
    if (d.small.ensure(sizeof(double) * ((size_t)nt + 2 + extraDoubles))) return FGP_ERR_ALLOC;
    std::vector<double> sc(nt);
    kernel_scales(ker, nt, theta, sc.data());
    *dScale = (double*)d.small.p;
    *dInfo = (int*)(*dScale + nt);
    if (dExtra) *dExtra = *dScale + nt + 2;
    CK(cudaMemcpyAsync(*dScale, sc.data(), sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(*dInfo, 0, sizeof(int), st));
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ preMats
extern "C" int fgp_premats(fgp_problem* P, int ker, double nugget, double sig2, const double* theta, double* L_out,
                           double* LInvY_out) {
    if (!P || !theta || ker < 1 || ker > 3 || !(sig2 > 0.0)) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    ProblemDev& pd = P->pd[0];
    const int n = P->n;
    int rc = ensure_factor_storage(P);
    if (rc) return rc;
    const long ld = pd.ldL;
    cudaStream_t st = d.sPanel[0];
    double* dScale; int* dInfo;
    if ((rc = stage_scales(P, d, ker, theta, 0, &dScale, &dInfo, nullptr, st))) return rc;
    AsmArgs a = fgp_train_asm(P, pd, ker, dScale);
    a.mult = sig2; a.diag_rel = nugget; a.M = pd.Lfac; a.ld = ld; a.strideM = 0;
    launch_assemble(a, 1, st);
    launch_fill_yrow(pd.Lfac, ld, 0, 1, n, pd.y, n, st);
    TrapDesc t{};
    t.M = pd.Lfac; t.ld = ld; t.strideM = 0; t.batch = 1;
    t.ncols = n; t.nrows = n + 1; t.holeLo = n + 1; t.holeHi = n + 1; t.colHoleLo = n; t.colHoleHi = n;
    t.cpre = 0; t.upperRows = 0; t.W = pd.Wall; t.strideW = 0; t.info = dInfo;
    if (!(ctx->dag && !ctx->profile && chol_dag_supported(n, n + 1) && !d.dagFlags.ensure(chol_dag_flag_bytes(n, 1)) &&
          launch_chol_dag(t.M, t.ld, t.strideM, 1, n, t.W, t.strideW, t.info, d.dagFlags.p, st, 1, ctx->dag_order))) {
        rc = trap_factor(ctx, 0, t, st, d.sBulk, ctx->lookahead != 0);
        if (rc) return rc;
    }
    FGP_COUNT_WORK(2, 1);
    int hinfo = 0;
    CK(cudaMemcpyAsync(&hinfo, dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(wait_stream(ctx, d, st));
    CK(cudaGetLastError());
    if (hinfo != 0) { P->hasModel = false; P->loocvValid = false; return hinfo; }
    set_model(P, ker, nugget, sig2, theta);
    return fetch_premats(P, L_out, LInvY_out);
}

// ------------------------------------------------------------------------------------------------ (de)serialisation
extern "C" int fgp_problem_restore(fgp_problem* P, int ker, double nugget, double sig2, const double* theta, const double* L,
                                   const double* LInvY) {
    if (!P || !theta || !L || !LInvY || ker < 1 || ker > 3 || !(sig2 > 0.0)) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    ProblemDev& pd = P->pd[0];
    const int n = P->n;
    for (int i = 0; i < n; ++i)
        if (!(L[(size_t)i * n + i] > 0.0)) { fgp_set_err(ctx, "fgp_problem_restore: L must have a positive diagonal"); return FGP_ERR_ARG; }
    int rc = ensure_factor_storage(P);
    if (rc) return rc;
    cudaStream_t st = d.sPanel[0];
    // the stored factor (lower triangle) and L^-1 y go up as they are; inv(L_kk) of the diagonal blocks is rebuilt
    rc = copy_matrix(ctx, d, false, const_cast<double*>(L), n, pd.Lfac, pd.ldL, n, n, 1, st);
    if (rc) return rc;
    if (d.small.ensure(sizeof(double) * (size_t)n)) return FGP_ERR_ALLOC;
    CK(cudaMemcpyAsync(d.small.p, LInvY, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    launch_fill_yrow(pd.Lfac, pd.ldL, 0, 1, n, (const double*)d.small.p, n, st);
    launch_trtri_blocks(pd.Lfac, pd.ldL, n, pd.Wall, st);
    CK(wait_stream(ctx, d, st));
    CK(cudaGetLastError());
    set_model(P, ker, nugget, sig2, theta);
    return 0;
}

// ------------------------------------------------------------------------------------------------ predict
extern "C" int fgp_predict(fgp_problem* P, int m, const double* sNew, const double* const* coefsNew, int detail,
                           double* mean, double* sd, double* Ktp, double* Kpp) {
    if (!P || m <= 0 || !mean || !sd) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    if (!P->hasModel) { fgp_set_err(ctx, "fgp_predict: call fgp_premats first"); return FGP_ERR_STATE; }
    if ((P->ds > 0 && !sNew) || (P->df > 0 && !coefsNew)) return FGP_ERR_ARG;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    ProblemDev& pd = P->pd[0];
    const int n = P->n, nt = P->nterms;
    double* Xnew = nullptr;
    int rc = upload_new(P, d, m, sNew, coefsNew, &Xnew);
    if (rc) return rc;
    const int holeHi = (int)roundup(n + 1, FGP_TILE);
    const long ld = pd.ldL;
    const int cap = (int)std::min<long>(ld - holeHi, m);
    if (cap < 1) { fgp_set_err(ctx, "fgp_predict: stored factor has no free rows"); return FGP_ERR_STATE; }
    cudaStream_t st = d.sPanel[0];
    double *dScale, *dRes; int* dInfo;
    if ((rc = stage_scales(P, d, P->ker, P->theta.data(), 2 * (size_t)m, &dScale, &dInfo, &dRes, st))) return rc;
    double* dMean = dRes; double* dNorm = dRes + m;
    AsmArgs a{};
    a.XQ = pd.X; a.nQ = n; a.ldQ = n; a.ldP = m;
    a.units = pd.units; a.nunits = (int)P->units.size(); a.scale = dScale; a.nterms = nt; a.ker = P->ker; a.sym = 0;
    a.mult = P->sig2; a.M = pd.Lfac; a.ld = ld; a.strideM = 0; a.rowOff = holeHi; a.colOff = 0;
    for (int c0 = 0; c0 < m; c0 += cap) {
        const int mc = std::min(cap, m - c0);
        // K(new, train) = sig2 * R(new, train) into the free rows under the factor     (R/4_prediction_SF.R:6)
        a.XP = Xnew + c0; a.nP = mc;
        launch_assemble(a, 1, st);
        TrapDesc t{};
        t.M = pd.Lfac; t.ld = ld; t.strideM = 0; t.batch = 1;
        t.ncols = n; t.nrows = holeHi + mc; t.holeLo = n; t.holeHi = holeHi; t.colHoleLo = n; t.colHoleHi = n;
        t.cpre = cdiv(n, FGP_TILE); t.upperRows = 0; t.W = pd.Wall; t.strideW = 0; t.info = dInfo;
        // solve mode of the dataflow kernel (one launch), else the launch-per-step schedule with every tile prefactored
        if (!(ctx->dag && !ctx->profile && !d.dagFlags.ensure(chol_dag_solve_flag_bytes(n, mc)) &&
              launch_chol_dag_solve(pd.Lfac, ld, n, holeHi, mc, pd.Wall, d.dagFlags.p, st))) {
            rc = trap_factor(ctx, 0, t, st, d.sBulk, ctx->lookahead != 0);
            if (rc) return rc;
        }
        // mean = LInvK' LInvY ; |LInvK_i|^2 for sd     (R/4_prediction_SF.R:10-11)
        launch_rowstats(pd.Lfac, ld, 0, 1, holeHi, mc, n, 0, pd.Lfac + n, ld, 0, dNorm + c0, dMean + c0, st);
    }
    std::vector<double> hn(m);
    CK(cudaMemcpyAsync(mean, dMean, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hn.data(), dNorm, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    CK(wait_stream(ctx, d, st));
    CK(cudaGetLastError());
    for (int i = 0; i < m; ++i) sd[i] = sqrt(std::max(P->sig2 - hn[i], 0.0));
    if (detail == FGP_DETAIL_FULL) {
        // K.tp (n x m) and K.pp (m x m) are assembled in column chunks that fit the workspace and streamed out
        a.M = nullptr; a.rowOff = 0; a.colOff = 0;
        const size_t budget = (size_t)256 << 20;
        if (Ktp) {
            const int wc = (int)std::max<size_t>(1, std::min<size_t>((size_t)m, budget / (sizeof(double) * (size_t)n)));
            if (d.work.ensure(sizeof(double) * (size_t)n * wc)) return FGP_ERR_ALLOC;
            for (int c0 = 0; c0 < m; c0 += wc) {
                const int mc = std::min(wc, m - c0);
                AsmArgs b = a;
                b.XP = pd.X; b.nP = n; b.ldP = n; b.XQ = Xnew + c0; b.nQ = mc; b.ldQ = m;
                b.M = (double*)d.work.p; b.ld = n;
                launch_assemble(b, 1, st);
                if ((rc = copy_matrix(ctx, d, true, Ktp + (size_t)c0 * n, n, (double*)d.work.p, n, n, mc, 0, st))) return rc;
            }
        }
        if (Kpp) {   // sig2 * (R_pp + nugget I)      (R/4_prediction_SF.R:16-17)
            if (d.work.ensure(sizeof(double) * (size_t)m * m)) return FGP_ERR_ALLOC;
            AsmArgs b = a;
            b.XP = Xnew; b.nP = m; b.ldP = m; b.XQ = Xnew; b.nQ = m; b.ldQ = m; b.sym = 1; b.mirror = 1;
            b.diag_rel = P->nugget; b.M = (double*)d.work.p; b.ld = m;
            launch_assemble(b, 1, st);
            if ((rc = copy_matrix(ctx, d, true, Kpp, m, (double*)d.work.p, m, m, m, 0, st))) return rc;
        }
        CK(cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ simulate
// private-copy path (more simulation points than free rows under the factor): one trapezoid with the new columns
static int simulate_copy(fgp_problem* P, int m, double* Xnew, int nsim, const double* Z, double nugget_sm, double* sims,
                         double* mean, double* sd) {
    fgp_ctx* ctx = P->ctx;
    DevState& d = ctx->dev[0];
    ProblemDev& pd = P->pd[0];
    const int n = P->n, nt = P->nterms;
    const int holeHi = (int)roundup(n, FGP_TILE);
    const int nrows = holeHi + m, ncols = holeHi + m;
    const long ld = pick_ld(nrows);
    const size_t trapBytes = sizeof(double) * (size_t)ld * ncols;
    const size_t extra = sizeof(double) * (2 * (size_t)m + (size_t)m * nsim * 2);
    if (d.work.ensure(trapBytes + extra)) return FGP_ERR_ALLOC;
    double* Mw = (double*)d.work.p;
    double* dMean = (double*)((char*)d.work.p + trapBytes);
    double* dNorm = dMean + m;
    double* dZ = dNorm + m;
    double* dSims = dZ + (size_t)m * nsim;
    const int ncTall = cdiv(ncols, FGP_TILE), ncTA = cdiv(n, FGP_TILE);
    if (d.wbuf.ensure(sizeof(double) * (size_t)ncTall * FGP_TILE * FGP_TILE)) return FGP_ERR_ALLOC;
    double* Ww = (double*)d.wbuf.p;
    cudaStream_t st = d.sPanel[0];
    double* dScale; int* dInfo;
    int rc = stage_scales(P, d, P->ker, P->theta.data(), 0, &dScale, &dInfo, nullptr, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dZ, Z, sizeof(double) * (size_t)m * nsim, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpy2DAsync(Mw, sizeof(double) * ld, pd.Lfac, sizeof(double) * pd.ldL, sizeof(double) * n, n,
                         cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(Ww, pd.Wall, sizeof(double) * (size_t)ncTA * FGP_TILE * FGP_TILE, cudaMemcpyDeviceToDevice, st));
    AsmArgs a{};
    a.XP = Xnew; a.nP = m; a.ldP = m; a.XQ = pd.X; a.nQ = n; a.ldQ = n;
    a.units = pd.units; a.nunits = (int)P->units.size(); a.scale = dScale; a.nterms = nt; a.ker = P->ker; a.sym = 0;
    a.mult = P->sig2; a.M = Mw; a.ld = ld; a.strideM = 0; a.rowOff = holeHi; a.colOff = 0;
    launch_assemble(a, 1, st);                       // K.ts'  (R/5_simulation_SF.R:8)
    AsmArgs b = a;                                   // K.ss + nugget.sm I  (R/5_simulation_SF.R:9,14)
    b.XQ = Xnew; b.nQ = m; b.ldQ = m; b.sym = 1; b.diag_abs = nugget_sm; b.colOff = holeHi;
    launch_assemble(b, 1, st);
    TrapDesc t{};
    t.M = Mw; t.ld = ld; t.strideM = 0; t.batch = 1;
    t.ncols = ncols; t.nrows = nrows; t.holeLo = n; t.holeHi = holeHi; t.colHoleLo = n; t.colHoleHi = holeHi;
    t.cpre = ncTA; t.upperRows = 0; t.W = Ww; t.strideW = 0; t.info = dInfo;
    rc = trap_factor(ctx, 0, t, st, d.sBulk, ctx->lookahead != 0);
    if (rc) return rc;
    launch_rowstats(Mw, ld, 0, 1, holeHi, m, n, 0, pd.Lfac + n, pd.ldL, 0, dNorm, dMean, st);
    launch_lower_times(Mw + holeHi + (size_t)holeHi * ld, ld, m, dZ, nsim, dMean, dSims, st);
    int hinfo = 0;
    std::vector<double> hn(m), hm(m);
    CK(cudaMemcpyAsync(&hinfo, dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hm.data(), dMean, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hn.data(), dNorm, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sims, dSims, sizeof(double) * (size_t)m * nsim, cudaMemcpyDeviceToHost, st));
    CK(wait_stream(ctx, d, st));
    CK(cudaGetLastError());
    if (hinfo != 0) return hinfo;    // chol(K.ss - ...) failed: base R's error (quirk Q5)
    if (mean) memcpy(mean, hm.data(), sizeof(double) * m);
    if (sd) for (int i = 0; i < m; ++i) sd[i] = sqrt(std::max(P->sig2 - hn[i], 0.0));
    return 0;
}

extern "C" int fgp_simulate(fgp_problem* P, int m, const double* sNew, const double* const* coefsNew, int nsim,
                            const double* Z, double nugget_sm, double* sims, double* mean, double* sd) {
    if (!P || m <= 0 || nsim <= 0 || !Z || !sims) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    if (!P->hasModel) { fgp_set_err(ctx, "fgp_simulate: call fgp_premats first"); return FGP_ERR_STATE; }
    if ((P->ds > 0 && !sNew) || (P->df > 0 && !coefsNew)) return FGP_ERR_ARG;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    ProblemDev& pd = P->pd[0];
    const int n = P->n, nt = P->nterms;
    double* Xnew = nullptr;
    int rc = upload_new(P, d, m, sNew, coefsNew, &Xnew);
    if (rc) return rc;
    const int holeHi = (int)roundup(n + 1, FGP_TILE);
    const long ld = pd.ldL;
    if ((long)holeHi + m > ld) return simulate_copy(P, m, Xnew, nsim, Z, nugget_sm, sims, mean, sd);
    // ---- in place: LInvK rows under the factor, conditional covariance and its factor in the workspace
    const long ldc = pick_ld(m), ldz = roundup(nsim, 16);
    const int mT = cdiv(m, FGP_TILE);
    const size_t cBytes = sizeof(double) * (size_t)ldc * m, zBytes = sizeof(double) * (size_t)ldz * m;
    const size_t sBytes = sizeof(double) * (size_t)nsim * m;
    if (d.work.ensure(cBytes + zBytes + sBytes + sizeof(double) * 2 * (size_t)m)) return FGP_ERR_ALLOC;
    if (d.wbuf.ensure(sizeof(double) * (size_t)mT * FGP_TILE * FGP_TILE)) return FGP_ERR_ALLOC;
    double* Css = (double*)d.work.p;
    double* dZt = (double*)((char*)d.work.p + cBytes);
    double* dSims = (double*)((char*)dZt + zBytes);
    double* dMean = (double*)((char*)dSims + sBytes);
    double* dNorm = dMean + m;
    cudaStream_t st = d.sPanel[0];
    double* dScale; int* dInfo;
    if ((rc = stage_scales(P, d, P->ker, P->theta.data(), 0, &dScale, &dInfo, nullptr, st))) return rc;
    // -Z' (nsim x m): the noise term Lc Z becomes a C -= A B' product with A = -Z', B = Lc     (R/5_simulation_SF.R:14-15)
    std::vector<double> zt((size_t)ldz * m, 0.0);
    for (int j = 0; j < m; ++j)
        for (int s = 0; s < nsim; ++s) zt[(size_t)s + (size_t)j * ldz] = -Z[(size_t)j + (size_t)s * m];
    CK(cudaMemcpyAsync(dZt, zt.data(), zBytes, cudaMemcpyHostToDevice, st));
    AsmArgs a{};
    a.XP = Xnew; a.nP = m; a.ldP = m; a.XQ = pd.X; a.nQ = n; a.ldQ = n;
    a.units = pd.units; a.nunits = (int)P->units.size(); a.scale = dScale; a.nterms = nt; a.ker = P->ker; a.sym = 0;
    a.mult = P->sig2; a.M = pd.Lfac; a.ld = ld; a.strideM = 0; a.rowOff = holeHi; a.colOff = 0;
    launch_assemble(a, 1, st);                       // K.ts'  (R/5_simulation_SF.R:8)
    AsmArgs b = a;                                   // K.ss + nugget.sm I  (R/5_simulation_SF.R:9,14)
    b.XQ = Xnew; b.nQ = m; b.ldQ = m; b.sym = 1; b.diag_abs = nugget_sm; b.M = Css; b.ld = ldc; b.rowOff = 0; b.colOff = 0;
    launch_assemble(b, 1, st);
    TrapDesc t{};
    t.M = pd.Lfac; t.ld = ld; t.strideM = 0; t.batch = 1;
    t.ncols = n; t.nrows = holeHi + m; t.holeLo = n; t.holeHi = holeHi; t.colHoleLo = n; t.colHoleHi = n;
    t.cpre = cdiv(n, FGP_TILE); t.upperRows = 0; t.W = pd.Wall; t.strideW = 0; t.info = dInfo;
    if (!(ctx->dag && !ctx->profile && !d.dagFlags.ensure(chol_dag_solve_flag_bytes(n, m)) &&
          launch_chol_dag_solve(pd.Lfac, ld, n, holeHi, m, pd.Wall, d.dagFlags.p, st))) {
        rc = trap_factor(ctx, 0, t, st, d.sBulk, ctx->lookahead != 0);  // LInvK' in the rows under the factor (:10)
        if (rc) return rc;
    }
    launch_rowstats(pd.Lfac, ld, 0, 1, holeHi, m, n, 0, pd.Lfac + n, ld, 0, dNorm, dMean, st);
    // K.ss - t(LInvK) LInvK: one symmetric update with K = n                                     (:14)
    launch_gemm_plain(pd.Lfac + holeHi, ld, pd.Lfac + holeHi, ld, Css, ldc, m, m, n, true, st);
    TrapDesc c{};
    c.M = Css; c.ld = ldc; c.strideM = 0; c.batch = 1;
    c.ncols = m; c.nrows = m; c.holeLo = m; c.holeHi = m; c.colHoleLo = m; c.colHoleHi = m;
    c.cpre = 0; c.upperRows = 0; c.W = (double*)d.wbuf.p; c.strideW = 0; c.info = dInfo;
    // chol of the conditional covariance: a plain m x m matrix (no y row)
    if (!(ctx->dag && !ctx->profile && !d.dagFlags.ensure(chol_dag_flag_bytes(m, 1)) &&
          launch_chol_dag(c.M, c.ld, 0, 1, m, c.W, 0, c.info, d.dagFlags.p, st, 1, ctx->dag_order, 0))) {
        rc = trap_factor(ctx, 0, c, st, d.sBulk, ctx->lookahead != 0);
        if (rc) return rc;
    }
    // sims = mean + (Lc Z)'  as  sims(s, i) = mean_i - sum_j (-Z')(s, j) Lc(i, j)
    launch_zero_strict_upper(Css, ldc, m, st);
    launch_bcast_rows(dSims, nsim, nsim, m, dMean, st);
    launch_gemm_plain(dZt, ldz, Css, ldc, dSims, nsim, nsim, m, m, false, st);
    int hinfo = 0;
    std::vector<double> hn(m), hm(m);
    CK(cudaMemcpyAsync(&hinfo, dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hm.data(), dMean, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hn.data(), dNorm, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    CK(wait_stream(ctx, d, st));
    CK(cudaGetLastError());
    if (hinfo != 0) return hinfo;    // chol(K.ss - ...) failed: base R's error (quirk Q5)
    if ((rc = copy_matrix(ctx, d, true, sims, nsim, dSims, nsim, nsim, m, 0, st))) return rc;
    if (mean) memcpy(mean, hm.data(), sizeof(double) * m);
    if (sd) for (int i = 0; i < m; ++i) sd[i] = sqrt(std::max(P->sig2 - hn[i], 0.0));
    return 0;
}

// ------------------------------------------------------------------------------------------------ LOOCV
extern "C" int fgp_loocv(fgp_problem* P, double* y_pre, double* q2) {
    if (!P || (!y_pre && !q2)) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    if (!P->hasModel) { fgp_set_err(ctx, "fgp_loocv: call fgp_premats first"); return FGP_ERR_STATE; }
    const int n = P->n;
    if (P->loocvValid) {
        // same model, asked again (getFitness, then plot(): R/7_plottingFunctions.R:9-11 recomputes solve(R))
        if (y_pre) memcpy(y_pre, P->loocvY.data(), sizeof(double) * n);
        if (q2) *q2 = P->loocvQ2;
        return 0;
    }
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    ProblemDev& pd = P->pd[0];
    const int holeHi = (int)roundup(n + 1, FGP_TILE);
    const int nrows = holeHi + n;
    const long ld = pick_ld(nrows);
    const int ncT = cdiv(n, FGP_TILE);
    const size_t trapBytes = sizeof(double) * (size_t)ld * n;
    if (d.work.ensure(trapBytes + sizeof(double) * 2 * (size_t)n)) return FGP_ERR_ALLOC;
    if (d.wbuf.ensure(sizeof(double) * (size_t)ncT * FGP_TILE * FGP_TILE)) return FGP_ERR_ALLOC;
    double* Mw = (double*)d.work.p;
    double* dDiag = (double*)((char*)d.work.p + trapBytes);
    double* dRy = dDiag + n;
    cudaStream_t st = d.sPanel[0];
    double* dScale; int* dInfo;
    int rc = stage_scales(P, d, P->ker, P->theta.data(), 0, &dScale, &dInfo, nullptr, st);
    if (rc) return rc;
    // R = tcrossprod(L)/sig2 + diag(nugget) = R0 + 2 nugget I  (quirk Q2, R/8_outilsStats.R:3)
    AsmArgs a = fgp_train_asm(P, pd, P->ker, dScale);
    a.diag_rel = 2.0 * P->nugget; a.M = Mw; a.ld = ld; a.strideM = 0;
    launch_assemble(a, 1, st);
    launch_fill_yrow(Mw, ld, 0, 1, n, pd.y, n, st);
    launch_fill_identity_rows(Mw, ld, 0, 1, holeHi, n, st);
    TrapDesc t{};
    t.M = Mw; t.ld = ld; t.strideM = 0; t.batch = 1;
    t.ncols = n; t.nrows = nrows; t.holeLo = n + 1; t.holeHi = holeHi; t.colHoleLo = n; t.colHoleHi = n;
    t.cpre = 0; t.upperRows = 1; t.W = (double*)d.wbuf.p; t.strideW = 0; t.info = dInfo;
    // dataflow kernel: factor the (n + 1)-row part, then the identity rows in its solve mode (two launches); else one
    // launch-per-step pass over the whole trapezoid
    bool dagDone = false;
    // (from 32 tile columns on: below, the two chain-bound launches lose to the single pass -- measured 0.90 vs 0.75 ms at
    // n = 1024, 1.80 vs 1.60 ms at 2048, 13.6 vs 18.2 ms at 8192)
    if (ctx->dag && !ctx->profile && ncT >= 32 &&
        !d.dagFlags.ensure(std::max(chol_dag_flag_bytes(n, 1), chol_dag_solve_flag_bytes(n, n)))) {
        if (launch_chol_dag(Mw, ld, 0, 1, n, t.W, 0, dInfo, d.dagFlags.p, st, 1, ctx->dag_order, 1))
            dagDone = launch_chol_dag_solve(Mw, ld, n, holeHi, n, t.W, d.dagFlags.p, st, 1);
        if (!dagDone) {         // (the factor may have run: start over from the assembled matrix)
            launch_assemble(a, 1, st);
            launch_fill_yrow(Mw, ld, 0, 1, n, pd.y, n, st);
            launch_fill_identity_rows(Mw, ld, 0, 1, holeHi, n, st);
        }
    }
    if (!dagDone) {
        rc = trap_factor(ctx, 0, t, st, d.sBulk, false);
        if (rc) return rc;
    }
    FGP_COUNT_WORK(1, 1);
    // X = L^-T in the extra rows: diag(Rinv)_i = |X_i.|^2 ; (Rinv y)_i = X_i. (L^-1 y)
    launch_rowstats(Mw, ld, 0, 1, holeHi, n, n, 1, Mw + n, ld, 0, dDiag, dRy, st);
    std::vector<double> hd(n), hr(n);
    int hinfo = 0;
    CK(cudaMemcpyAsync(&hinfo, dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hd.data(), dDiag, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hr.data(), dRy, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CK(wait_stream(ctx, d, st));
    CK(cudaGetLastError());
    if (hinfo != 0) return hinfo;
    // y_pre = y - Rinv y / diag(Rinv)  (R/8_outilsStats.R:6);  Q2 = 1 - mean((y-yhat)^2)/mean((y-ybar)^2)
    double ybar = 0.0;
    for (int i = 0; i < n; ++i) ybar += P->y[i];
    ybar /= n;
    double se = 0.0, st2 = 0.0;
    P->loocvY.resize(n);
    for (int i = 0; i < n; ++i) {
        const double yh = P->y[i] - hr[i] / hd[i];
        P->loocvY[i] = yh;
        se += (P->y[i] - yh) * (P->y[i] - yh);
        st2 += (P->y[i] - ybar) * (P->y[i] - ybar);
    }
    P->loocvQ2 = 1.0 - (se / n) / (st2 / n);
    P->loocvValid = true;
    if (y_pre) memcpy(y_pre, P->loocvY.data(), sizeof(double) * n);
    if (q2) *q2 = P->loocvQ2;
    return 0;
}

// ------------------------------------------------------------------------------------------------ update() fast paths
// sigma^2 substitution (upd_subHypers, R/6_updating.R:510-524 re-assembles and re-factors): K' = (s'/s) K, so
// L' = sqrt(s'/s) L, (L^-1 y)' = (L^-1 y) / sqrt(s'/s) -- one pass over the stored factor.
extern "C" int fgp_update_var(fgp_problem* P, double sig2_new, double* L_out, double* LInvY_out) {
    if (!P || !(sig2_new > 0.0)) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    if (!P->hasModel) { fgp_set_err(ctx, "fgp_update_var: the problem has no model"); return FGP_ERR_STATE; }
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    ProblemDev& pd = P->pd[0];
    const int n = P->n;
    const double a = sqrt(sig2_new / P->sig2);
    cudaStream_t st = d.sPanel[0];
    launch_scale_factor(pd.Lfac, pd.ldL, n, pd.Wall, (long)cdiv(n, FGP_TILE) * FGP_TILE * FGP_TILE, a, st);
    CK(wait_stream(ctx, d, st));
    CK(cudaGetLastError());
    P->sig2 = sig2_new;          // leave-one-out predictions do not depend on sigma^2: the cache stays valid
    return fetch_premats(P, L_out, LInvY_out);
}

// output substitution (upd_subData with sOut.sb only, R/6_updating.R:154-160, refits): the factor does not depend on y;
// only L^-1 y is solved again -- the new y goes through the prefactored tiles as one extra row, O(n^2).
extern "C" int fgp_update_y(fgp_problem* P, const double* y_new, double* LInvY_out) {
    if (!P || !y_new) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    if (!P->hasModel) { fgp_set_err(ctx, "fgp_update_y: the problem has no model"); return FGP_ERR_STATE; }
    int rc = fgp_problem_set_y(P, y_new);        // clears hasModel: the model is re-established below
    if (rc) return rc;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    ProblemDev& pd = P->pd[0];
    const int n = P->n;
    const int holeHi = (int)roundup(n + 1, FGP_TILE);
    cudaStream_t st = d.sPanel[0];
    double* dScale; int* dInfo;
    if ((rc = stage_scales(P, d, P->ker, P->theta.data(), 0, &dScale, &dInfo, nullptr, st))) return rc;
    launch_fill_yrow(pd.Lfac, pd.ldL, 0, 1, holeHi, pd.y, n, st);
    TrapDesc t{};
    t.M = pd.Lfac; t.ld = pd.ldL; t.strideM = 0; t.batch = 1;
    t.ncols = n; t.nrows = holeHi + 1; t.holeLo = n; t.holeHi = holeHi; t.colHoleLo = n; t.colHoleHi = n;
    t.cpre = cdiv(n, FGP_TILE); t.upperRows = 0; t.W = pd.Wall; t.strideW = 0; t.info = dInfo;
    // the one row through the solve mode of the dataflow kernel (from 16 column tiles on), else launch by launch
    if (!(ctx->dag && !ctx->profile && !d.dagFlags.ensure(chol_dag_solve_flag_bytes(n, 1)) &&
          launch_chol_dag_solve(pd.Lfac, pd.ldL, n, holeHi, 1, pd.Wall, d.dagFlags.p, st))) {
        rc = trap_factor(ctx, 0, t, st, d.sBulk, false);
        if (rc) return rc;
    }
    launch_copy_row(pd.Lfac, pd.ldL, n, holeHi, n, st);
    CK(wait_stream(ctx, d, st));
    CK(cudaGetLastError());
    P->hasModel = true; P->loocvValid = false;
    return fetch_premats(P, nullptr, LInvY_out);
}

// data addition / deletion / substitution with retained hyper-parameters (upd_add, upd_del, upd_subData,
// R/6_updating.R:7-89, 91-323, 325-500: every branch refits with fgpm()): prob_new holds the updated data set, prob_old
// the fitted model.  The leading points the two share (same coordinates, same order) keep their block of the factor:
// with n0 = that count rounded down to a multiple of 128, L[0:n0, 0:n0] and inv(L_kk) are copied, everything from row n0
// on is assembled afresh and eliminated against the prefactored tiles (the engine simulate uses), and only the
// trailing (n' - n0)^2 block is factored.  Appending k points to n costs ~ n^2 (k + n mod 128) instead of n^3 / 3.
extern "C" int fgp_premats_reuse(fgp_problem* Pn, const fgp_problem* Po, double* L_out, double* LInvY_out, int* n_reused) {
    if (!Pn || !Po) return FGP_ERR_ARG;
    fgp_ctx* ctx = Pn->ctx;
    if (!Po->hasModel || Po->ctx != ctx) { fgp_set_err(ctx, "fgp_premats_reuse: prob_old must hold a model of the same context"); return FGP_ERR_STATE; }
    if (Pn->ds != Po->ds || Pn->df != Po->df || Pn->p != Po->p || Pn->dis != Po->dis || Pn->ncoord != Po->ncoord || Pn->Jw != Po->Jw) {
        fgp_set_err(ctx, "fgp_premats_reuse: the two problems differ in structure (inputs, projection dimensions, distances or Gram matrices)");
        return FGP_ERR_ARG;
    }
    const int n = Pn->n, no = Po->n, nc = Pn->ncoord;
    // common leading points: identical coordinates row by row
    int keep = 0;
    for (; keep < std::min(n, no); ++keep) {
        bool same = true;
        for (int c = 0; c < nc && same; ++c) same = Pn->X[(size_t)c * n + keep] == Po->X[(size_t)c * no + keep];
        if (!same) break;
    }
    const int n0 = (keep / FGP_TILE) * FGP_TILE;
    if (n_reused) *n_reused = n0;
    if (n0 == 0) return fgp_premats(Pn, Po->ker, Po->nugget, Po->sig2, Po->theta.data(), L_out, LInvY_out);
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    int rc = ensure_factor_storage(Pn);
    if (rc) return rc;
    ProblemDev& pd = Pn->pd[0];
    const ProblemDev& po = Po->pd[0];
    const long ld = pd.ldL;
    const int nt = Pn->nterms, cpre = n0 / FGP_TILE;
    cudaStream_t st = d.sPanel[0];
    double* dScale; int* dInfo;
    if ((rc = stage_scales(Pn, d, Po->ker, Po->theta.data(), 0, &dScale, &dInfo, nullptr, st))) return rc;
    CK(cudaMemcpy2DAsync(pd.Lfac, sizeof(double) * ld, po.Lfac, sizeof(double) * po.ldL, sizeof(double) * n0, n0, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(pd.Wall, po.Wall, sizeof(double) * (size_t)cpre * FGP_TILE * FGP_TILE, cudaMemcpyDeviceToDevice, st));
    // rows n0.. : K(rest, kept) as a rectangle, K(rest, rest) + nugget as a symmetric block
    AsmArgs a{};
    a.XP = pd.X + n0; a.nP = n - n0; a.ldP = n; a.XQ = pd.X; a.nQ = n0; a.ldQ = n;
    a.units = pd.units; a.nunits = (int)Pn->units.size(); a.scale = dScale; a.nterms = nt; a.ker = Po->ker; a.sym = 0;
    a.mult = Po->sig2; a.M = pd.Lfac; a.ld = ld; a.strideM = 0; a.rowOff = n0; a.colOff = 0;
    if (n > n0) {
        launch_assemble(a, 1, st);
        AsmArgs b = a;
        b.XQ = pd.X + n0; b.nQ = n - n0; b.sym = 1; b.diag_rel = Po->nugget; b.colOff = n0;
        launch_assemble(b, 1, st);
    }
    launch_fill_yrow(pd.Lfac, ld, 0, 1, n, pd.y, n, st);
    TrapDesc t{};
    t.M = pd.Lfac; t.ld = ld; t.strideM = 0; t.batch = 1;
    t.ncols = n; t.nrows = n + 1; t.holeLo = n0; t.holeHi = n0; t.colHoleLo = n; t.colHoleHi = n;
    t.cpre = cpre; t.upperRows = 0; t.W = pd.Wall; t.strideW = 0; t.info = dInfo;
    rc = trap_factor(ctx, 0, t, st, d.sBulk, ctx->lookahead != 0);
    if (rc) return rc;
    int hinfo = 0;
    CK(cudaMemcpyAsync(&hinfo, dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(wait_stream(ctx, d, st));
    CK(cudaGetLastError());
    if (hinfo != 0) { Pn->hasModel = false; return hinfo; }
    set_model(Pn, Po->ker, Po->nugget, Po->sig2, Po->theta.data());
    return fetch_premats(Pn, L_out, LInvY_out);
}
