// Context / problem objects behind the C ABI (include/fungp_b200.h).
#pragma once
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <memory>

struct FgpTimers {
    enum { ASSEMBLY = 0, PANEL = 1, UPDATE = 2, FACTOR = 3, TRSM = 4, NCLASS = 5 };
    struct Rec { cudaEvent_t a, b; int cls; double flops; };
    std::vector<cudaEvent_t> pool;
    std::vector<Rec> recs;
    size_t used = 0;
    cudaEvent_t cur = nullptr;
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    void reset() { used = 0; recs.clear(); }
    void begin(cudaStream_t st) { cur = get(); cudaEventRecord(cur, st); }
    void end(cudaStream_t st, int cls, double flops) {
        cudaEvent_t e = get();
        cudaEventRecord(e, st);
        recs.push_back({cur, e, cls, flops});
    }
    // sums per class: ms, count, flops
    void collect(double ms[NCLASS], double cnt[NCLASS], double fl[NCLASS]) {
        for (int i = 0; i < NCLASS; ++i) ms[i] = cnt[i] = fl[i] = 0.0;
        for (auto& r : recs) {
            cudaEventSynchronize(r.b);
            float t = 0.f;
            cudaEventElapsedTime(&t, r.a, r.b);
            ms[r.cls] += t; cnt[r.cls] += 1.0; fl[r.cls] += r.flops;
        }
    }
    FgpTimers() = default;
    FgpTimers(const FgpTimers&) = delete;             // owns its events: movable, not copyable
    FgpTimers& operator=(const FgpTimers&) = delete;
    FgpTimers(FgpTimers&& o) noexcept { swap(o); }
    FgpTimers& operator=(FgpTimers&& o) noexcept { swap(o); return *this; }
    void swap(FgpTimers& o) { pool.swap(o.pool); recs.swap(o.recs); std::swap(used, o.used); std::swap(cur, o.cur); }
    ~FgpTimers() { for (auto e : pool) cudaEventDestroy(e); }
};

// growable device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// caching device allocator (per context and device): cudaFree synchronises the whole device, which stalls every other
// worker context of fgp_fit_batch; problems are created and destroyed per candidate, so their blocks are recycled
struct DevPool {
    struct Blk { void* p; size_t cap; bool used; };
    std::vector<Blk> blks;
    void* get(size_t bytes) {
        bytes = std::max<size_t>(bytes, 256);
        int best = -1;
        for (int i = 0; i < (int)blks.size(); ++i)
            if (!blks[i].used && blks[i].cap >= bytes && (best < 0 || blks[i].cap < blks[best].cap)) best = i;
        if (best >= 0 && blks[best].cap <= 2 * bytes + 4096) { blks[best].used = true; return blks[best].p; }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        blks.push_back({p, bytes, true});
        return p;
    }
    void put(void* p) {
        for (auto& b : blks) if (b.p == p) { b.used = false; return; }
    }
    void release() { for (auto& b : blks) cudaFree(b.p); blks.clear(); }
};

struct DevState {
    int device = 0;
    static const int MAXSTREAMS = 8;
    cudaStream_t sPanel[MAXSTREAMS] = {nullptr};
    cudaStream_t sBulk = nullptr;
    cudaEvent_t evA = nullptr, evB = nullptr, evJoin[MAXSTREAMS] = {nullptr}, evMark[4] = {nullptr};
    cudaEvent_t evWait = nullptr;   // blocking-sync event of wait_stream (fit_batch workers)
    cudaEvent_t evDag[2] = {nullptr, nullptr};   // around the last dataflow-Cholesky launch of fgp_nll_batch (fgp_last_factor_ms)
    bool dagTimed = false;
    DevBuf work;      // factorisation workspaces (batch of trapezoids)
    DevBuf wbuf;      // inverse diagonal blocks
    DevBuf small;     // thetas / scales / results
    DevBuf aux;       // new-point coordinates etc.
    DevBuf dagFlags;  // tile-ready flags and the task counter of the dataflow Cholesky (chol_dag.cu)
    DevPool pool;     // per-problem blocks (coordinates, y, stored factor)
    void* pinned = nullptr; size_t pinnedCap = 0;
    FgpTimers timers;
    double lastAsmMs = 0, lastFactorMs = 0;
    long launches = 0;
};

struct fgp_ctx {
    std::vector<DevState> dev;
    std::string err;
    int profile = 0;
    int lookahead = 1;
    int streams = 8;
    int fit_workers = 8;      // fgp_fit_batch, fit_mode 0: concurrent candidate fits (host threads) per GPU
    int fit_mode = 1;         // 1: lock-step -- one host thread per GPU advances all in-flight candidates together (fitlock.cu); 0: one thread per candidate
    int gemm_tma = 1;         // operands of the DMMA GEMM by TMA (gemm_tma.cu); 0 = cp.async kernel
    int host_threads = 0;     // host threads of the chunked matrix transfers; 0 = min(16, cores)
    int out_prezeroed = 0;    // 1: the caller's L_out matrices are zero-filled on entry (calloc / numpy.zeros): only the lower triangle is written
    int host_thp = 1;         // ask for transparent huge pages on large host output matrices (model.cu copy_matrix)
    int dag_order = 1;        // dataflow Cholesky: 1 = tiles of a column claimed matrix by matrix, 0 = row by row over the matrices
    int dag_group = 1;        // dataflow Cholesky: tile columns claimed together (1, 2, 4; see chol_dag.cu)
    int dag = 1;              // likelihood batches, preMats and the lock-step candidate driver factor with the persistent dataflow
                              // kernel (chol_dag.cu) when n is a multiple of 128; 0: always the launch-per-step engine (chol.cu)
    int gemm_flat = 8;        // TMA kernel: persistent CTAs over all (tile, element) pairs from this many pairs per SM on (0 = never)
    int pdl = 1;              // programmatic dependent launch along the factorisation's kernel chain
    int fused_assembly = 1;   // batched likelihood calls assemble the matrices of one structure with the batch-fused kernel
    int fit_points = 0;       // lock-step: likelihood points per round and lane (0 = by memory, at most 512)
    int fit_lanes = 0;        // lock-step: rounds in flight per GPU (0 = 3 below n = 2048, else 1)
    int predict_reserve = 2048;   // rows kept free under the stored factor so that predict / simulate solve in place
    int graphs = 0;           // option "graphs": replay small-n likelihood calls as CUDA graphs (second call onwards)
    int blockingSync = 0;     // host waits sleep instead of spinning (set on fit_batch worker contexts)
    int fit_blocking = -1;    // option "fit_blocking": worker contexts of fgp_fit_batch use blocking waits (-1 = auto)
    int outer = 16;           // column tiles per outer panel for batches (trailing update K = outer*128)
    int la_r8 = 44, la_r4 = 28;   // adaptive look-ahead widths: 8 tiles while more than la_r8 remain, 4 while more than la_r4 (chol.cu)
    int outer_single = 0;     // same for a single factorisation with look-ahead; 0 = by size (chol.cu)
    double timing[8] = {0};
    std::mutex mu;
    // fgp_fit_batch: worker contexts (one per device and worker slot), created on first use and kept: creating streams
    // and workspaces per call cost more than a candidate at n = 1024
    std::vector<fgp_ctx*> fitWorkers;
    // lock-step driver state of a worker context (streams, pinned staging, workspaces), released with the context
    void* lockState = nullptr;
    void (*lockFree)(void*) = nullptr;
};

static inline long fgp_roundup(long a, long b) { return (a + b - 1) / b * b; }
// leading dimension: multiple of 16 doubles (128 B); avoid large power-of-two strides
static inline long fgp_pick_ld(long rows) {
    long ld = fgp_roundup(rows, 16);
    if (ld % 512 == 0) ld += 16;
    return ld;
}
// host wait for a stream honouring the context's waiting policy (spin / blocking event / poll-and-yield), api.cu
cudaError_t fgp_wait_stream(fgp_ctx* ctx, DevState& d, cudaStream_t st);

// a captured launch sequence of one batched likelihood call (small n: one stream, ~26 launches), see nll_range
struct NllGraph {
    int ker = 0, B = 0;
    double nugget = 0;
    const void* work = nullptr; const void* wbuf = nullptr; const void* small = nullptr; const void* pinned = nullptr;
    int outer = 0;
    const void* dagFlags = nullptr; int dag = 0;      // the captured sequence contains the dataflow kernel with this flag buffer
    cudaGraphExec_t exec = nullptr;     // null: seen once, not captured yet
};

struct ProblemDev {
    bool ready = false;
    double* X = nullptr;      // n x ncoord coordinate matrix (column-major, ld = n)
    FgpUnit* units = nullptr;
    double* y = nullptr;
    // model state after fgp_premats (device 0 only)
    double* Lfac = nullptr;   // trapezoid with the factor of sig2 (R + nugget I) and the (L^-1 y)' row
    long ldL = 0;
    double* Wall = nullptr;   // inverse diagonal blocks of Lfac
    std::vector<NllGraph> graphs;
};

struct fgp_problem {
    fgp_ctx* ctx = nullptr;
    int n = 0, ds = 0, df = 0;
    std::vector<int> p, dis;
    int nterms = 0;           // d
    int ncoord = 0;           // columns of the coordinate matrix
    std::vector<FgpUnit> units;
    std::vector<int> termFirstUnit;     // nterms + 1
    std::vector<double> X;    // host copy (replicated lazily to the other devices)
    std::vector<double> y;
    std::vector<std::vector<double>> J;   // per functional input (p x p), as given
    std::vector<std::vector<double>> Jw;  // per functional input: W with J = W W' (L2_bygroup; empty otherwise), whitens new points too
    std::vector<ProblemDev> pd;
    // model state
    bool hasModel = false;
    int ker = 0;
    double nugget = 0, sig2 = 0;
    std::vector<double> theta;
    // leave-one-out predictions of the current model, kept for the callers that ask again (getFitness, then the
    // calibration plot: R/7_plottingFunctions.R:9-11, 569-572 recompute solve(R) for the same model)
    bool loocvValid = false;
    std::vector<double> loocvY;
    double loocvQ2 = 0.0;
};

int problem_upload(fgp_problem* P, int slot);
// build the coordinate matrix rows for m points (host): scalars, byindex coefs, (c, cJ) pairs for bygroup
void build_coords(const fgp_problem* P, int m, const double* sIn, const double* const* coefs, std::vector<double>& X);
void kernel_scales(int ker, int nterms, const double* theta, double* scale);
AsmArgs fgp_train_asm(const fgp_problem* P, const ProblemDev& pd, int ker, const double* scale_dev);
