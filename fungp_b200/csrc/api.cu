// C ABI of libfungp_b200.so (include/fungp_b200.h): contexts, problems and the likelihood / prediction /
// simulation / LOOCV entry points.  No CPU fallback: every numeric path launches the sm_100a kernels.
#include "ctx.h"
#include <sched.h>
#include "../../include/fungp_b200.h"
#include <cstring>
#include <thread>

// ------------------------------------------------------------------------------------------------ errors
void fgp_set_err(fgp_ctx* ctx, const std::string& msg) {
    if (!ctx) return;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err = msg;
}
void fgp_set_err_cuda(fgp_ctx* ctx, const char* what, cudaError_t e, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    fgp_set_err(ctx, buf);
    cudaGetLastError();
}
#define CK(expr)                                                              \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            fgp_set_err_cuda(ctx, #expr, _e, __FILE__, __LINE__);             \
            return FGP_ERR_CUDA;                                              \
        }                                                                     \
    } while (0)

// Host wait for a stream.  Worker contexts of fgp_fit_batch (blockingSync) sleep on an event created with
// cudaEventBlockingSync instead of spinning: 8 workers per GPU on an 8-GPU box are 64 waiting threads, more than the
// box has cores.  User contexts keep the lower-latency spin of cudaStreamSynchronize.
cudaError_t fgp_wait_stream(fgp_ctx* ctx, DevState& d, cudaStream_t st) {
    if (!ctx->blockingSync) return cudaStreamSynchronize(st);
    if (!d.evWait) {
        cudaError_t e = cudaEventCreateWithFlags(&d.evWait, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(d.evWait, st);
    if (e != cudaSuccess) return e;
    if (ctx->blockingSync == 2) {
        // poll and yield: next to no added latency while a core is free, and waiting threads give way when there are
        // more of them than cores
        for (;;) {
            e = cudaEventQuery(d.evWait);
            if (e != cudaErrorNotReady) return e;
            sched_yield();
        }
    }
    return cudaEventSynchronize(d.evWait);
}

static inline cudaError_t wait_stream(fgp_ctx* ctx, DevState& d, cudaStream_t st) { return fgp_wait_stream(ctx, d, st); }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long roundup(long a, long b) { return fgp_roundup(a, b); }
static inline long pick_ld(long rows) { return fgp_pick_ld(rows); }

std::atomic<long long> g_fgp_launches{0};
std::atomic<long long> g_fgp_work[4];
extern int g_plain_tma;

extern "C" const char* fgp_version(void) { return "fungp_b200 0.1 (sm_100a, FP64 DMMA)"; }

// ------------------------------------------------------------------------------------------------ context
extern "C" int fgp_ctx_create(int n_devices, const int* device_ids, fgp_ctx** out) {
    if (!out) return FGP_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return FGP_ERR_NODEVICE;   // no CPU fallback by design
    }
    if (n_devices <= 0) n_devices = count;
    if (n_devices > FGP_MAXDEV) n_devices = FGP_MAXDEV;
    fgp_ctx* ctx = new fgp_ctx();
    ctx->dev.resize(n_devices);
    for (int i = 0; i < n_devices; ++i) {
        DevState& d = ctx->dev[i];
        d.device = device_ids ? device_ids[i] : i;
        if (d.device < 0 || d.device >= count) { delete ctx; return FGP_ERR_ARG; }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d.device) != cudaSuccess || prop.major < 10) {
            // the kernels are sm_100a only
            delete ctx;
            return FGP_ERR_NODEVICE;
        }
        if (cudaSetDevice(d.device) != cudaSuccess) { delete ctx; return FGP_ERR_CUDA; }
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        for (int s = 0; s < DevState::MAXSTREAMS; ++s) {
            cudaStreamCreateWithPriority(&d.sPanel[s], cudaStreamNonBlocking, hi);   // panel = critical path
            cudaEventCreateWithFlags(&d.evJoin[s], cudaEventDisableTiming);
        }
        cudaStreamCreateWithPriority(&d.sBulk, cudaStreamNonBlocking, lo);
        cudaEventCreateWithFlags(&d.evA, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&d.evB, cudaEventDisableTiming);
    }
    cudaSetDevice(ctx->dev[0].device);
    *out = ctx;
    return 0;
}

extern "C" void fgp_ctx_destroy(fgp_ctx* ctx) {
    if (!ctx) return;
    for (fgp_ctx* w : ctx->fitWorkers) fgp_ctx_destroy(w);
    ctx->fitWorkers.clear();
    if (ctx->lockState && ctx->lockFree) {
        if (!ctx->dev.empty()) cudaSetDevice(ctx->dev[0].device);
        ctx->lockFree(ctx->lockState);
        ctx->lockState = nullptr;
    }
    for (auto& d : ctx->dev) {
        cudaSetDevice(d.device);
        cudaDeviceSynchronize();
        for (int s = 0; s < DevState::MAXSTREAMS; ++s) {
            if (d.sPanel[s]) cudaStreamDestroy(d.sPanel[s]);
            if (d.evJoin[s]) cudaEventDestroy(d.evJoin[s]);
        }
        if (d.sBulk) cudaStreamDestroy(d.sBulk);
        if (d.evA) cudaEventDestroy(d.evA);
        if (d.evB) cudaEventDestroy(d.evB);
        if (d.evWait) cudaEventDestroy(d.evWait);
        for (int e = 0; e < 2; ++e) if (d.evDag[e]) cudaEventDestroy(d.evDag[e]);
        d.work.release(); d.wbuf.release(); d.small.release(); d.aux.release(); d.dagFlags.release(); d.pool.release();
        if (d.pinned) cudaFreeHost(d.pinned);
    }
    delete ctx;
}

extern "C" int fgp_ctx_device_count(const fgp_ctx* ctx) { return ctx ? (int)ctx->dev.size() : 0; }
extern "C" const char* fgp_last_error(const fgp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int fgp_set_option(fgp_ctx* ctx, const char* name, double value) {
    if (!ctx || !name) return FGP_ERR_ARG;
    if (!strcmp(name, "profile")) ctx->profile = value != 0;
    else if (!strcmp(name, "lookahead")) ctx->lookahead = value != 0;
    else if (!strcmp(name, "fit_blocking")) ctx->fit_blocking = value < 0 ? -1 : std::min((int)value, 2);
    else if (!strcmp(name, "blocking_sync")) ctx->blockingSync = std::max(0, std::min((int)value, 2));
    else if (!strcmp(name, "graphs")) ctx->graphs = value != 0;
    else if (!strcmp(name, "fit_workers")) ctx->fit_workers = std::max(1, std::min((int)value, 16));
    else if (!strcmp(name, "fit_mode")) ctx->fit_mode = value != 0 ? 1 : 0;
    else if (!strcmp(name, "la_r8")) ctx->la_r8 = std::max(0, (int)value);
    else if (!strcmp(name, "la_r4")) ctx->la_r4 = std::max(0, (int)value);
    else if (!strcmp(name, "pdl")) ctx->pdl = value != 0 ? 1 : 0;
    else if (!strcmp(name, "host_threads")) ctx->host_threads = std::max(0, std::min((int)value, 64));
    else if (!strcmp(name, "out_prezeroed")) ctx->out_prezeroed = value != 0 ? 1 : 0;
    else if (!strcmp(name, "host_thp")) ctx->host_thp = value != 0 ? 1 : 0;
    else if (!strcmp(name, "dag")) ctx->dag = value != 0 ? 1 : 0;
    else if (!strcmp(name, "dag_order")) ctx->dag_order = value != 0 ? 1 : 0;
    else if (!strcmp(name, "dag_group")) ctx->dag_group = std::max(1, std::min((int)value, 16));
    else if (!strcmp(name, "gemm_flat")) ctx->gemm_flat = value < 0 ? 0 : (int)value;
    else if (!strcmp(name, "gemm_tma")) { ctx->gemm_tma = value != 0 ? 1 : 0; g_plain_tma = ctx->gemm_tma; }
    else if (!strcmp(name, "fused_assembly")) ctx->fused_assembly = value != 0 ? 1 : 0;
    else if (!strcmp(name, "fit_points")) ctx->fit_points = std::max(0, std::min((int)value, 4096));
    else if (!strcmp(name, "fit_lanes")) ctx->fit_lanes = std::max(0, std::min((int)value, 4));
    else if (!strcmp(name, "predict_reserve")) ctx->predict_reserve = std::max(128, std::min((int)value, 1 << 20));
    else if (!strcmp(name, "outer")) ctx->outer = std::max(1, std::min((int)value, 16));
    else if (!strcmp(name, "outer_single")) ctx->outer_single = std::max(0, std::min((int)value, 16));
    else if (!strcmp(name, "streams")) ctx->streams = std::max(1, std::min((int)value, (int)DevState::MAXSTREAMS));
    else { fgp_set_err(ctx, std::string("unknown option ") + name); return FGP_ERR_ARG; }
    return 0;
}

extern "C" int fgp_get_timing(fgp_ctx* ctx, double* out8) {
    if (!ctx || !out8) return FGP_ERR_ARG;
    DevState& d = ctx->dev[0];
    cudaSetDevice(d.device);
    double ms[FgpTimers::NCLASS], cnt[FgpTimers::NCLASS], fl[FgpTimers::NCLASS];
    d.timers.collect(ms, cnt, fl);
    out8[0] = ms[FgpTimers::ASSEMBLY];
    out8[1] = ms[FgpTimers::FACTOR];
    out8[2] = ms[FgpTimers::UPDATE];
    out8[3] = cnt[FgpTimers::UPDATE];
    out8[4] = fl[FgpTimers::UPDATE];
    out8[5] = ms[FgpTimers::PANEL];
    out8[6] = ms[FgpTimers::TRSM];
    out8[7] = cnt[FgpTimers::UPDATE] + cnt[FgpTimers::PANEL] + cnt[FgpTimers::TRSM] + cnt[FgpTimers::ASSEMBLY];
    return 0;
}

extern "C" int fgp_last_factor_ms(fgp_ctx* ctx, double* ms) {
    if (!ctx || !ms) return FGP_ERR_ARG;
    DevState& d = ctx->dev[0];
    *ms = 0.0;
    if (!d.dagTimed || !d.evDag[0] || !d.evDag[1]) return 0;
    CK(cudaSetDevice(d.device));
    float f = 0.f;
    if (cudaEventElapsedTime(&f, d.evDag[0], d.evDag[1]) != cudaSuccess) { cudaGetLastError(); return 0; }
    *ms = f;
    return 0;
}

extern "C" long long fgp_launch_count(void) { return g_fgp_launches.load(); }
extern "C" long long fgp_work_count(int which) { return (which >= 0 && which < 4) ? g_fgp_work[which].load() : -1; }

// device-side stopwatch over ALL of the library's streams on device 0: mark(i) records an event that follows every
// stream's queued work; elapsed(i, j) is the CUDA-event time between two marks.
extern "C" int fgp_mark(fgp_ctx* ctx, int idx) {
    if (!ctx || idx < 0 || idx >= 4) return FGP_ERR_ARG;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    if (!d.evMark[idx]) CK(cudaEventCreate(&d.evMark[idx]));
    for (int s = 0; s < DevState::MAXSTREAMS; ++s) {
        CK(cudaEventRecord(d.evJoin[s], d.sPanel[s]));
        CK(cudaStreamWaitEvent(d.sBulk, d.evJoin[s], 0));
    }
    CK(cudaEventRecord(d.evMark[idx], d.sBulk));
    return 0;
}
extern "C" int fgp_elapsed_ms(fgp_ctx* ctx, int i, int j, double* ms) {
    if (!ctx || !ms || i < 0 || j < 0 || i >= 4 || j >= 4) return FGP_ERR_ARG;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    if (!d.evMark[i] || !d.evMark[j]) return FGP_ERR_STATE;
    CK(cudaEventSynchronize(d.evMark[j]));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, d.evMark[i], d.evMark[j]));
    *ms = t;
    return 0;
}

extern "C" int fgp_measure_peaks(fgp_ctx* ctx, double* dmma_tflops, double* dfma_tflops, double* copy_gbs) {
    if (!ctx) return FGP_ERR_ARG;
    cudaSetDevice(ctx->dev[0].device);
    return fgp_peaks_device(dmma_tflops, dfma_tflops, copy_gbs);
}

// ------------------------------------------------------------------------------------------------ problem
// lower factor W (p x p, column-major) with J = W W': Cholesky; if J is not numerically positive definite (rank-
// deficient basis), the symmetric eigen-decomposition with negative eigenvalues clamped (the reference clamps the
// resulting squared distances at 0 instead, R/7_distanceFunctions.R:44)
void whitening_factor(const double* J, int p, std::vector<double>& W) {
    W.assign((size_t)p * p, 0.0);
    bool ok = true;
    for (int j = 0; j < p && ok; ++j) {
        double d = J[j + (size_t)j * p];
        for (int q = 0; q < j; ++q) d -= W[j + (size_t)q * p] * W[j + (size_t)q * p];
        if (!(d > 0.0) || !std::isfinite(d)) { ok = false; break; }
        const double l = sqrt(d);
        W[j + (size_t)j * p] = l;
        for (int i = j + 1; i < p; ++i) {
            double v = 0.5 * (J[i + (size_t)j * p] + J[j + (size_t)i * p]);
            for (int q = 0; q < j; ++q) v -= W[i + (size_t)q * p] * W[j + (size_t)q * p];
            W[i + (size_t)j * p] = v / l;
        }
    }
    if (ok) return;
    // cyclic Jacobi on the symmetrised J
    std::vector<double> A((size_t)p * p), V((size_t)p * p, 0.0);
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < p; ++j) A[i + (size_t)j * p] = 0.5 * (J[i + (size_t)j * p] + J[j + (size_t)i * p]);
    for (int i = 0; i < p; ++i) V[i + (size_t)i * p] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int j = 0; j < p; ++j)
            for (int i = 0; i < p; ++i) (i == j ? dg : off) += A[i + (size_t)j * p] * A[i + (size_t)j * p];
        if (off <= 1e-32 * dg || off == 0.0) break;
        for (int a = 0; a < p - 1; ++a)
            for (int b = a + 1; b < p; ++b) {
                const double apq = A[a + (size_t)b * p];
                if (apq == 0.0) continue;
                const double th = (A[b + (size_t)b * p] - A[a + (size_t)a * p]) / (2.0 * apq);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int r = 0; r < p; ++r) {
                    const double x = A[r + (size_t)a * p], y = A[r + (size_t)b * p];
                    A[r + (size_t)a * p] = c * x - sn * y; A[r + (size_t)b * p] = sn * x + c * y;
                }
                for (int r = 0; r < p; ++r) {
                    const double x = A[a + (size_t)r * p], y = A[b + (size_t)r * p];
                    A[a + (size_t)r * p] = c * x - sn * y; A[b + (size_t)r * p] = sn * x + c * y;
                }
                for (int r = 0; r < p; ++r) {
                    const double x = V[r + (size_t)a * p], y = V[r + (size_t)b * p];
                    V[r + (size_t)a * p] = c * x - sn * y; V[r + (size_t)b * p] = sn * x + c * y;
                }
            }
    }
    for (int j = 0; j < p; ++j) {
        const double w = sqrt(std::max(0.0, A[j + (size_t)j * p]));
        for (int i = 0; i < p; ++i) W[i + (size_t)j * p] = V[i + (size_t)j * p] * w;
    }
}

// whitened coefficients z = c W, J = W W': ||z_i - z_j||^2 = (c_i - c_j)' J (c_i - c_j)
// (the bilinear form of R/7_distanceFunctions.R:39-45); z must be zero-initialised
void whiten_columns(const double* c, int m, int p, const double* W, double* zcols) {
    for (int a = 0; a < p; ++a) {
        double* z = zcols + (size_t)a * m;
        for (int b = 0; b < p; ++b) {
            const double wba = W[b + (size_t)a * p];
            if (wba == 0.0) continue;
            const double* cb = c + (size_t)b * m;
            for (int r = 0; r < m; ++r) z[r] += cb[r] * wba;
        }
    }
}

void build_coords(const fgp_problem* P, int m, const double* sIn, const double* const* coefs, std::vector<double>& X) {
    X.assign((size_t)m * P->ncoord, 0.0);
    int col = 0;
    for (int s = 0; s < P->ds; ++s, ++col) memcpy(&X[(size_t)col * m], sIn + (size_t)s * m, sizeof(double) * m);
    for (int i = 0; i < P->df; ++i) {
        const int p = P->p[i];
        const double* c = coefs[i];
        if (P->dis[i] == FGP_DIST_BYINDEX) {
            for (int a = 0; a < p; ++a, ++col) memcpy(&X[(size_t)col * m], c + (size_t)a * m, sizeof(double) * m);
        } else {
            whiten_columns(c, m, p, P->Jw[i].data(), &X[(size_t)col * m]);
            col += p;
        }
    }
}

void kernel_scales(int ker, int nterms, const double* theta, double* scale) {
    const double c = ker == FGP_KER_GAUSS ? 1.0 : (ker == FGP_KER_MATERN52 ? sqrt(5.0) : sqrt(3.0));
    for (int t = 0; t < nterms; ++t) scale[t] = c / theta[t];
}

int problem_upload(fgp_problem* P, int slot) {
    fgp_ctx* ctx = P->ctx;
    ProblemDev& pd = P->pd[slot];
    if (pd.ready) return 0;
    CK(cudaSetDevice(ctx->dev[slot].device));
    DevPool& pool = ctx->dev[slot].pool;
    pd.X = (double*)pool.get(sizeof(double) * std::max<size_t>(1, P->X.size()));
    pd.units = (FgpUnit*)pool.get(sizeof(FgpUnit) * std::max<size_t>(1, P->units.size()));
    pd.y = (double*)pool.get(sizeof(double) * P->n);
    if (!pd.X || !pd.units || !pd.y) { fgp_set_err(ctx, "out of device memory"); return FGP_ERR_ALLOC; }
    // stream-ordered + synchronised: the consumers run on non-blocking streams that do not wait for the default stream
    cudaStream_t up = ctx->dev[slot].sPanel[0];
    CK(cudaMemcpyAsync(pd.X, P->X.data(), sizeof(double) * P->X.size(), cudaMemcpyHostToDevice, up));
    CK(cudaMemcpyAsync(pd.units, P->units.data(), sizeof(FgpUnit) * P->units.size(), cudaMemcpyHostToDevice, up));
    CK(cudaMemcpyAsync(pd.y, P->y.data(), sizeof(double) * P->n, cudaMemcpyHostToDevice, up));
    CK(wait_stream(ctx, ctx->dev[slot], up));
    pd.ready = true;
    return 0;
}

extern "C" int fgp_problem_create(fgp_ctx* ctx, int n, int ds, const double* sIn, int df, const int* p,
                                  const int* dis_type, const double* const* coefs, const double* const* J,
                                  const double* y, fgp_problem** out) {
    if (!ctx || !out || n <= 0 || ds < 0 || df < 0 || (ds == 0 && df == 0) || !y) return FGP_ERR_ARG;
    if ((ds > 0 && !sIn) || (df > 0 && (!p || !dis_type || !coefs || !J))) return FGP_ERR_ARG;
    *out = nullptr;
    std::unique_ptr<fgp_problem> P(new fgp_problem());
    P->ctx = ctx; P->n = n; P->ds = ds; P->df = df;
    int col = 0, term = 0;
    for (int s = 0; s < ds; ++s) {
        P->termFirstUnit.push_back((int)P->units.size());
        P->units.push_back({0, col++, term++, 1});
    }
    for (int i = 0; i < df; ++i) {
        if (p[i] <= 0 || (dis_type[i] != FGP_DIST_BYGROUP && dis_type[i] != FGP_DIST_BYINDEX)) {
            fgp_set_err(ctx, "fgp_problem_create: p[i] must be the number of coefficient columns (> 0) and dis_type 1|2");
            return FGP_ERR_ARG;
        }
        P->p.push_back(p[i]); P->dis.push_back(dis_type[i]);
        P->J.emplace_back(J[i], J[i] + (size_t)p[i] * p[i]);
        P->Jw.emplace_back();
        if (dis_type[i] == FGP_DIST_BYGROUP) whitening_factor(J[i], p[i], P->Jw.back());
        if (dis_type[i] == FGP_DIST_BYINDEX) {
            for (int a = 0; a < p[i]; ++a) {
                P->termFirstUnit.push_back((int)P->units.size());
                P->units.push_back({0, col++, term++, 1});
            }
        } else {
            P->termFirstUnit.push_back((int)P->units.size());
            for (int a = 0; a < p[i]; ++a) P->units.push_back({1, col++, term, a == p[i] - 1});
            ++term;
        }
    }
    P->termFirstUnit.push_back((int)P->units.size());
    P->nterms = term; P->ncoord = col;
    build_coords(P.get(), n, sIn, coefs, P->X);
    P->y.assign(y, y + n);
    P->pd.resize(ctx->dev.size());
    int rc = problem_upload(P.get(), 0);
    if (rc) return rc;
    *out = P.release();
    return 0;
}

extern "C" void fgp_problem_destroy(fgp_problem* P) {
    if (!P) return;
    for (size_t s = 0; s < P->pd.size(); ++s) {
        ProblemDev& pd = P->pd[s];
        if (!pd.ready && !pd.Lfac) continue;
        DevState& d = P->ctx->dev[s];
        cudaSetDevice(d.device);
        // every call on this problem has synchronised its streams before returning: the blocks go back to the pool
        if (pd.X) d.pool.put(pd.X);
        if (pd.units) d.pool.put(pd.units);
        if (pd.y) d.pool.put(pd.y);
        if (pd.Lfac) d.pool.put(pd.Lfac);
        if (pd.Wall) d.pool.put(pd.Wall);
        for (auto& gph : pd.graphs) if (gph.exec) cudaGraphExecDestroy(gph.exec);
    }
    delete P;
}

extern "C" int fgp_problem_n(const fgp_problem* P) { return P ? P->n : 0; }
extern "C" int fgp_problem_nls(const fgp_problem* P) { return P ? P->nterms : 0; }

extern "C" int fgp_problem_set_y(fgp_problem* P, const double* y) {
    if (!P || !y) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    P->y.assign(y, y + P->n);
    for (size_t s = 0; s < P->pd.size(); ++s)
        if (P->pd[s].ready) {
            CK(cudaSetDevice(ctx->dev[s].device));
            CK(cudaMemcpyAsync(P->pd[s].y, y, sizeof(double) * P->n, cudaMemcpyHostToDevice, ctx->dev[s].sPanel[0]));
            CK(wait_stream(ctx, ctx->dev[s], ctx->dev[s].sPanel[0]));
        }
    P->hasModel = false;
    P->loocvValid = false;
    return 0;
}

// ------------------------------------------------------------------------------------------------ distances
extern "C" int fgp_dist_max(fgp_problem* P, double* max_out) {
    if (!P || !max_out) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    if (d.small.ensure(sizeof(double) * std::max(1, P->nterms))) return FGP_ERR_ALLOC;
    double* dmax = (double*)d.small.p;
    CK(cudaMemsetAsync(dmax, 0, sizeof(double) * P->nterms, d.sPanel[0]));
    const int n = P->n;
    for (int t = 0; t < P->nterms; ++t) {
        const int u0 = P->termFirstUnit[t], u1 = P->termFirstUnit[t + 1];
        if (P->units[u0].kind == 1)
            launch_distmax(P->pd[0].X, n, n, P->pd[0].units, u0, u1, dmax + t, d.sPanel[0]);
    }
    std::vector<double> h(P->nterms);
    CK(cudaMemcpyAsync(h.data(), dmax, sizeof(double) * P->nterms, cudaMemcpyDeviceToHost, d.sPanel[0]));
    CK(wait_stream(ctx, d, d.sPanel[0]));
    for (int t = 0; t < P->nterms; ++t) {
        const FgpUnit& u = P->units[P->termFirstUnit[t]];
        if (u.kind == 1) max_out[t] = h[t];
        else {
            // max |x_i - x_j| = max - min of the column (abs(outer(-)), R/7_distanceFunctions.R:31)
            const double* x = &P->X[(size_t)u.col * n];
            double lo = x[0], hi = x[0];
            for (int i = 1; i < n; ++i) { lo = std::min(lo, x[i]); hi = std::max(hi, x[i]); }
            max_out[t] = hi - lo;
        }
    }
    return 0;
}

extern "C" int fgp_dist_materialize(fgp_problem* P, int l, double* out) {
    if (!P || !out || l < 0 || l >= P->nterms) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    const int n = P->n;
    if (d.work.ensure(sizeof(double) * (size_t)n * n)) return FGP_ERR_ALLOC;
    launch_distmat(P->pd[0].X, n, n, P->pd[0].units, P->termFirstUnit[l], P->termFirstUnit[l + 1], (double*)d.work.p, n,
                   d.sPanel[0]);
    CK(cudaMemcpyAsync(out, d.work.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, d.sPanel[0]));
    CK(wait_stream(ctx, d, d.sPanel[0]));
    return 0;
}

AsmArgs fgp_train_asm(const fgp_problem* P, const ProblemDev& pd, int ker, const double* scale_dev) {
    AsmArgs a{};
    a.XP = pd.X; a.nP = P->n; a.ldP = P->n;
    a.XQ = pd.X; a.nQ = P->n; a.ldQ = P->n;
    a.units = pd.units; a.nunits = (int)P->units.size();
    a.scale = scale_dev; a.nterms = P->nterms; a.ker = ker; a.sym = 1;
    a.mult = 1.0; a.diag_rel = 0.0; a.diag_abs = 0.0;
    a.rowOff = 0; a.colOff = 0; a.mirror = 0;
    return a;
}

extern "C" int fgp_assemble(fgp_problem* P, int ker, double nugget, const double* theta, double* R_out) {
    if (!P || !theta || !R_out || ker < 1 || ker > 3) return FGP_ERR_ARG;
    fgp_ctx* ctx = P->ctx;
    DevState& d = ctx->dev[0];
    CK(cudaSetDevice(d.device));
    const int n = P->n;
    if (d.work.ensure(sizeof(double) * (size_t)n * n)) return FGP_ERR_ALLOC;
    if (d.small.ensure(sizeof(double) * P->nterms)) return FGP_ERR_ALLOC;
    std::vector<double> sc(P->nterms);
    kernel_scales(ker, P->nterms, theta, sc.data());
    cudaStream_t st = d.sPanel[0];
    CK(cudaMemcpyAsync(d.small.p, sc.data(), sizeof(double) * P->nterms, cudaMemcpyHostToDevice, st));
    AsmArgs a = fgp_train_asm(P, P->pd[0], ker, (const double*)d.small.p);
    a.diag_rel = nugget; a.M = (double*)d.work.p; a.ld = n; a.strideM = 0; a.mirror = 1;
    launch_assemble(a, 1, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(R_out, d.work.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, st));
    CK(wait_stream(ctx, d, st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ likelihood
// one device's share of a batch: points [b0, b1)
static int nll_range(fgp_problem* P, int slot, int ker, double nugget, int b0, int b1, const double* thetas,
                     const double* var_known, double* nll, double* sig2, int* info) {
    fgp_ctx* ctx = P->ctx;
    DevState& d = ctx->dev[slot];
    CK(cudaSetDevice(d.device));
    int rc = problem_upload(P, slot);
    if (rc) return rc;
    const ProblemDev& pd = P->pd[slot];
    const int n = P->n, nt = P->nterms, B = b1 - b0;
    if (B <= 0) return 0;
    const long ld = pick_ld(n + 1);
    const size_t matBytes = sizeof(double) * (size_t)ld * n;
    const int ncT = cdiv(n, FGP_TILE);
    const size_t wBytes = sizeof(double) * (size_t)ncT * FGP_TILE * FGP_TILE;
    int chunk = std::min(B, 4096);
    if (matBytes * chunk > d.work.cap || wBytes * chunk > d.wbuf.cap) {
        // the workspaces must grow: size the chunk against what the device has left (not on the steady-state path)
        size_t freeB = 0, totB = 0;
        cudaMemGetInfo(&freeB, &totB);
        // at most 60 % of what the device has left, and 64 GB (of 180): 7 full-storage matrices of n = 32768 per chunk
        const size_t budget = std::min<size_t>((size_t)((freeB + d.work.cap + d.wbuf.cap) * 0.6), (size_t)64 << 30);
        chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)chunk, budget / (matBytes + wBytes)));
    }
    if (d.work.ensure(matBytes * chunk) || d.wbuf.ensure(wBytes * chunk)) {
        fgp_set_err(ctx, "fgp_nll_batch: out of device memory");
        return FGP_ERR_ALLOC;
    }
    // small buffer: scales [chunk*nt] | nll [chunk] | sig2 [chunk] | info [chunk ints, padded] | var_known [1]
    const size_t resBytes = sizeof(double) * 2 * (size_t)chunk + sizeof(int) * (size_t)roundup(chunk, 2);
    const size_t smallBytes = sizeof(double) * ((size_t)chunk * nt + 1) + resBytes;
    if (d.small.ensure(smallBytes)) return FGP_ERR_ALLOC;
    double* dScale = (double*)d.small.p;
    double* dNll = dScale + (size_t)chunk * nt;
    double* dSig = dNll + chunk;
    int* dInfo = (int*)(dSig + chunk);
    double* dVar = (double*)((char*)dNll + resBytes);
    // pinned staging for the length-scale multipliers (H2D) and the results (one D2H per chunk): the copies are truly
    // asynchronous, so the only host wait of the call is wait_stream below
    const size_t pinBytes = sizeof(double) * (size_t)chunk * nt + resBytes;
    if (pinBytes > d.pinnedCap) {
        if (d.pinned) cudaFreeHost(d.pinned);
        d.pinned = nullptr; d.pinnedCap = 0;
        if (cudaHostAlloc(&d.pinned, std::max<size_t>(pinBytes, 1 << 16), cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            fgp_set_err(ctx, "fgp_nll_batch: out of pinned host memory");
            return FGP_ERR_ALLOC;
        }
        d.pinnedCap = std::max<size_t>(pinBytes, 1 << 16);
    }
    double* hScale = (double*)d.pinned;
    char* hRes = (char*)d.pinned + sizeof(double) * (size_t)chunk * nt;
    const bool prof = ctx->profile != 0;
    if (prof) d.timers.reset();
    const double hVar = var_known ? *var_known : 0.0;

    for (int c0 = 0; c0 < B; c0 += chunk) {
        const int cb = std::min(chunk, B - c0);
        FGP_COUNT_WORK(0, cb);
        for (int i = 0; i < cb; ++i) kernel_scales(ker, nt, thetas + (size_t)(b0 + c0 + i) * nt, hScale + (size_t)i * nt);
        // NOTE: every host->device copy below is issued on the stream that consumes it.  The library's streams are
        // non-blocking: they do not wait for the legacy default stream, and a cudaMemcpy from pageable memory may return
        // before its DMA has landed -- kernels on another stream would then read the previous batch's length-scales
        // (seen under concurrent fgp_fit_batch workers).
        // split the chunk over streams so that one group's panel phase overlaps another group's trailing update
        // small matrices: ONE batched launch sequence (gridDim.y = batch) -- more streams only add false dependencies in
        // the hardware queues (measured at n = 1024, 15 points: 1.3 ms with 1 stream, 3.4 ms with 8); large ones: several
        // groups so that one group's panel phase hides under another group's trailing update
        int ns = prof ? 1 : std::min(ctx->streams, cb);
        if (n < 2048) ns = 1;
        else if (n < 4096) ns = std::min(ns, 2);
        // dataflow Cholesky: the whole chunk in one persistent launch (one group, one stream)
        const bool useDag = ctx->dag && !prof && chol_dag_supported(n, n + 1) && !d.dagFlags.ensure(chol_dag_flag_bytes(n, cb));
        if (useDag) ns = 1;
        const int per = cdiv(cb, ns);
        // The points of a batch share every coordinate difference: with enough of them ONE batch-fused assembly launch
        // builds all matrices of the chunk (bit-identical entries) on the first stream, and the groups fork from there;
        // otherwise every group assembles its own matrices on its stream.
        const bool fuseAll = cb >= 4 && ctx->fused_assembly && assemble_fusable(nt, (int)P->units.size());
        auto enqueue_front = [&]() -> int {
            cudaStream_t st = d.sPanel[0];
            CK(cudaMemcpyAsync(dScale, hScale, sizeof(double) * (size_t)cb * nt, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(dInfo, 0, sizeof(int) * cb, st));
            if (var_known) CK(cudaMemcpyAsync(dVar, &hVar, sizeof(double), cudaMemcpyHostToDevice, st));
            AsmArgs a = fgp_train_asm(P, pd, ker, dScale);
            a.diag_rel = nugget; a.M = (double*)d.work.p; a.ld = ld; a.strideM = (long)ld * n;
            if (prof) d.timers.begin(st);
            launch_assemble_fused(a, nullptr, 1, cb, st);
            if (prof) d.timers.end(st, FgpTimers::ASSEMBLY, 0.0);
            if (ns > 1) {
                CK(cudaEventRecord(d.evA, st));
                for (int s = 1; s < ns; ++s) CK(cudaStreamWaitEvent(d.sPanel[s], d.evA, 0));
            }
            return 0;
        };
        // one group = one launch sequence on its stream
        auto enqueue_group = [&](int s) -> int {
            const int g0 = s * per, g1 = std::min(cb, g0 + per);
            if (g1 <= g0) return 0;
            const int gb = g1 - g0;
            cudaStream_t st = d.sPanel[s];
            double* Mg = (double*)d.work.p + (size_t)g0 * ld * n;
            if (!fuseAll) {
                CK(cudaMemcpyAsync(dScale + (size_t)g0 * nt, hScale + (size_t)g0 * nt, sizeof(double) * (size_t)gb * nt,
                                   cudaMemcpyHostToDevice, st));
                CK(cudaMemsetAsync(dInfo + g0, 0, sizeof(int) * gb, st));
                if (var_known) CK(cudaMemcpyAsync(dVar, &hVar, sizeof(double), cudaMemcpyHostToDevice, st));
                AsmArgs a = fgp_train_asm(P, pd, ker, dScale + (size_t)g0 * nt);
                a.diag_rel = nugget; a.M = Mg; a.ld = ld; a.strideM = (long)ld * n;
                if (prof) d.timers.begin(st);
                launch_assemble(a, gb, st);
                if (prof) d.timers.end(st, FgpTimers::ASSEMBLY, 0.0);
            }
            launch_fill_yrow(Mg, ld, (long)ld * n, gb, n, pd.y, n, st);
            TrapDesc t{};
            t.M = Mg; t.ld = ld; t.strideM = (long)ld * n; t.batch = gb;
            t.ncols = n; t.nrows = n + 1; t.holeLo = n + 1; t.holeHi = n + 1;
            t.colHoleLo = n; t.colHoleHi = n; t.cpre = 0; t.upperRows = 0;
            t.W = (double*)d.wbuf.p + (size_t)g0 * ncT * FGP_TILE * FGP_TILE; t.strideW = (long)ncT * FGP_TILE * FGP_TILE;
            t.info = dInfo + g0;
            if (prof) d.timers.begin(st);
            cudaEvent_t fa = prof ? d.timers.cur : nullptr;
            bool dagDone = false;
            if (useDag) {
                // timed with two events on its own stream (one launch: the price is two event records per call)
                for (int e = 0; e < 2; ++e)
                    if (!d.evDag[e]) CK(cudaEventCreate(&d.evDag[e]));
                CK(cudaEventRecord(d.evDag[0], st));
                dagDone = launch_chol_dag(t.M, t.ld, t.strideM, gb, n, t.W, t.strideW, t.info, d.dagFlags.p, st, ctx->dag_group, ctx->dag_order);
                CK(cudaEventRecord(d.evDag[1], st));
                d.dagTimed = dagDone;
            }
            if (!dagDone) {
                const int rcf = trap_factor(ctx, slot, t, st, d.sBulk, ctx->lookahead && ns == 1 && gb == 1);
                if (rcf) return rcf;
            }
            if (prof) { d.timers.cur = fa; d.timers.end(st, FgpTimers::FACTOR, 0.0); }
            launch_nll_reduce(Mg, ld, (long)ld * n, gb, n, n, var_known ? dVar : nullptr, dInfo + g0, dNll + g0, dSig + g0, st);
            return 0;
        };
        // Small matrices, option "graphs": the whole call (copy in, ~26 launches, copy out) is captured the second time
        // the same shape is seen on this problem and replayed from then on: one launch per call instead of ~30 driver
        // calls -- what the host threads of an 8-GPU candidate search spend most of their time in.
        NllGraph* gr = nullptr;
        bool replayed = false;
        if (ctx->graphs && ns == 1 && cb > 1 && !prof && !var_known && c0 == 0 && cb == B && n < 2048) {
            ProblemDev& pdw = P->pd[slot];
            for (auto& g : pdw.graphs)
                if (g.ker == ker && g.B == cb && g.nugget == nugget && g.work == d.work.p && g.wbuf == d.wbuf.p &&
                    g.small == d.small.p && g.pinned == d.pinned && g.outer == ctx->outer && g.dag == (int)useDag &&
                    g.dagFlags == (useDag ? d.dagFlags.p : nullptr)) { gr = &g; break; }
            if (!gr) {
                if (pdw.graphs.size() < 8) {
                    NllGraph g; g.ker = ker; g.B = cb; g.nugget = nugget; g.work = d.work.p; g.wbuf = d.wbuf.p;
                    g.small = d.small.p; g.pinned = d.pinned; g.outer = ctx->outer; g.dag = (int)useDag; g.dagFlags = useDag ? d.dagFlags.p : nullptr;
                    pdw.graphs.push_back(g);          // first sighting: run normally (also warms the launchers' statics)
                }
            } else {
                cudaStream_t st = d.sPanel[0];
                if (!gr->exec) {
                    cudaGraph_t graph = nullptr;
                    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                    int rce = fuseAll ? enqueue_front() : 0;
                    if (!rce) rce = enqueue_group(0);
                    cudaError_t ec = cudaMemcpyAsync(hRes, dNll, resBytes, cudaMemcpyDeviceToHost, st);
                    cudaError_t ee = cudaStreamEndCapture(st, &graph);
                    if (rce || ec != cudaSuccess || ee != cudaSuccess || !graph) {
                        if (graph) cudaGraphDestroy(graph);
                        cudaGetLastError();
                        gr->B = -1;                   // never try this entry again; fall through to the normal path
                        gr = nullptr;
                    } else {
                        cudaError_t ei = cudaGraphInstantiate(&gr->exec, graph, 0);
                        cudaGraphDestroy(graph);
                        if (ei != cudaSuccess) { cudaGetLastError(); gr->exec = nullptr; gr->B = -1; gr = nullptr; }
                    }
                }
                if (gr && gr->exec) {
                    CK(cudaGraphLaunch(gr->exec, st));
                    replayed = true;
                }
            }
        }
        if (!replayed) {
            if (fuseAll && (rc = enqueue_front())) return rc;
            for (int s = 0; s < ns; ++s) {
                rc = enqueue_group(s);
                if (rc) return rc;
            }
        }
        // join the groups on stream 0, fetch nll | sig2 | info with one copy, and wait once
        for (int s = 1; s < ns; ++s) {
            if (!d.evJoin[s]) CK(cudaEventCreateWithFlags(&d.evJoin[s], cudaEventDisableTiming));
            CK(cudaEventRecord(d.evJoin[s], d.sPanel[s]));
            CK(cudaStreamWaitEvent(d.sPanel[0], d.evJoin[s], 0));
        }
        if (!replayed) CK(cudaMemcpyAsync(hRes, dNll, resBytes, cudaMemcpyDeviceToHost, d.sPanel[0]));
        CK(wait_stream(ctx, d, d.sPanel[0]));
        CK(cudaGetLastError());
        memcpy(nll + b0 + c0, hRes, sizeof(double) * cb);
        memcpy(sig2 + b0 + c0, hRes + sizeof(double) * chunk, sizeof(double) * cb);
        memcpy(info + b0 + c0, hRes + sizeof(double) * 2 * chunk, sizeof(int) * cb);
    }
    return 0;
}

extern "C" int fgp_nll_batch(fgp_problem* P, int ker, double nugget, int B, const double* thetas,
                             const double* var_known, double* nll, double* sig2, int* info) {
    if (!P || B < 0 || !thetas || !nll || !sig2 || !info || ker < 1 || ker > 3) return FGP_ERR_ARG;
    if (B == 0) return 0;
    fgp_ctx* ctx = P->ctx;
    for (long i = 0; i < (long)B * P->nterms; ++i)
        if (!(thetas[i] > 0.0)) { fgp_set_err(ctx, "fgp_nll_batch: length-scales must be > 0"); return FGP_ERR_ARG; }
    const int nd = (int)ctx->dev.size();
    if (nd == 1 || B == 1) return nll_range(P, 0, ker, nugget, 0, B, thetas, var_known, nll, sig2, info);
    // independent points: contiguous shares per GPU, one host thread each, host-side gather (no collective)
    std::vector<int> rcs(nd, 0);
    std::vector<std::thread> th;
    for (int s = 0; s < nd; ++s) {
        const int b0 = (int)((long)B * s / nd), b1 = (int)((long)B * (s + 1) / nd);
        th.emplace_back([=, &rcs]() { rcs[s] = nll_range(P, s, ker, nugget, b0, b1, thetas, var_known, nll, sig2, info); });
    }
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->dev[0].device);
    for (int s = 0; s < nd; ++s) if (rcs[s]) return rcs[s];
    return 0;
}

