// Lock-step candidate scoring: ONE host thread per GPU advances the fits of all in-flight candidate
// structures together.  Reference per candidate: fitNtests_ACO (R/3_ant_admin.R:810-831) = formatSol_ACO (:301-373) ->
// fgpm(): setBounds / setSPoints / optimHypers (R/3_training_SF.R:57-223; optim L-BFGS-B with finite differences) ->
// getFitness (R/1_Xfgpm_Class.R:522-552).
//
// A candidate is a small state machine (presample -> L-BFGS-B starts -> value at the optimum -> score) whose optimiser
// runs in reverse communication (lbfgsb.hpp Machine).  Every ROUND the driver gathers what all machines of a lane need
// next -- 20 presample points here, the 2d+1 finite-difference points of an iterate there -- into ONE heterogeneous
// batch: each batch element carries its own unit table and length-scales (AsmElem), all share n, so assembly, the y
// row, the whole blocked Cholesky and the nll reduction are the same launches a single fgp_nll_batch makes, only
// hundreds of matrices wide instead of 2d+1.  One H2D blob in, one D2H of scalars out, one host wait per round and
// lane; three lanes per GPU alternate so that the device never waits for the host.  Leave-one-out scoring of the
// candidates that finished in a round is batched the same way (the 2n x n trapezoids of fgp_loocv side by side).
// Per-item results do not depend on what else is in the round (elements are independent, look-ahead is off, the
// schedule is fixed by n), hence not on lane sizes, lanes or the number of GPUs.
#include "fitcommon.h"
#include "lbfgsb.hpp"
#include <sched.h>
#include <cstring>
#include <thread>

using namespace fgp_fit;

namespace {

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

struct Cand {
    int item = -1;
    Decoded D;
    int d = 0, ker = 0, n = 0;
    std::vector<FgpUnit> units;        // columns = absolute columns of the device arena
    std::vector<double> lo, up, pts;
    enum Phase { PRESAMPLE, OPT, FINAL, SCORE } phase = PRESAMPLE;
    std::vector<fgp_lbfgsb::Machine> mach;
    int need = 0;                      // points this candidate may ask for in one round
    // the current round
    int slot0 = 0, nreq = 0, lslot = -1;
    std::vector<double> req;           // d x nreq length-scale vectors
    bool have = false;
    double bestf = 0.0, sig2 = 0.0;
    std::vector<double> bestx;
};

struct Lane {
    cudaStream_t st[DevState::MAXSTREAMS] = {nullptr};
    cudaEvent_t evFork = nullptr, evJoin[DevState::MAXSTREAMS] = {nullptr}, evDone = nullptr;
    DevBuf work, wbuf, blob, res, lwork, lwbuf, dag;     // dag: flags + task counter of the dataflow Cholesky, per lane
    char* hBlob = nullptr; size_t hBlobCap = 0;
    char* hRes = nullptr; size_t hResCap = 0;
    std::vector<Cand*> cands;
    bool busy = false;
    int Btot = 0, Lcnt = 0;
    int ensure_pinned(char** p, size_t* cap, size_t bytes) {
        if (bytes <= *cap) return 0;
        if (*p) cudaFreeHost(*p);
        *p = nullptr; *cap = 0;
        const size_t want = std::max<size_t>(bytes * 2, 1 << 16);
        if (cudaHostAlloc((void**)p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return -1; }
        *cap = want;
        return 0;
    }
};

// device-resident coordinate columns shared by all candidates of a call on one GPU
struct Arena {
    int rows = 0;
    long capCols = 0, used = 0;
    DevBuf buf, tmp;
    std::map<std::tuple<int, int, int, int, int>, long> col0;     // (input or -1-scalar, p, basis, dist, rep) -> first column
    std::vector<double*> dY;                                      // output vector per replicate (index rep + 1; 0 = full data)
    DevBuf ybuf;
};

struct LockState {
    Lane lane[4];
    Arena arena;
    bool streamsReady = false;
};

void lock_free(void* p) {
    LockState* L = (LockState*)p;
    for (Lane& ln : L->lane) {
        for (int s = 0; s < DevState::MAXSTREAMS; ++s) {
            if (ln.st[s]) cudaStreamDestroy(ln.st[s]);
            if (ln.evJoin[s]) cudaEventDestroy(ln.evJoin[s]);
        }
        if (ln.evFork) cudaEventDestroy(ln.evFork);
        if (ln.evDone) cudaEventDestroy(ln.evDone);
        ln.work.release(); ln.wbuf.release(); ln.blob.release(); ln.res.release(); ln.lwork.release(); ln.lwbuf.release(); ln.dag.release();
        if (ln.hBlob) cudaFreeHost(ln.hBlob);
        if (ln.hRes) cudaFreeHost(ln.hRes);
    }
    L->arena.buf.release(); L->arena.tmp.release(); L->arena.ybuf.release();
    delete L;
}

#define LCK(expr)                                                             \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            fgp_set_err_cuda(wc, #expr, _e, __FILE__, __LINE__);              \
            return FGP_ERR_CUDA;                                              \
        }                                                                     \
    } while (0)

struct Driver {
    fgp_ctx* parent;
    fgp_ctx* wc;            // worker context of this GPU (its own streams and pools)
    Shared& S;
    const std::vector<int>& order;
    LockState* L = nullptr;
    int n = 0;              // rows of every item (n, or the training rows of a hold-out split)
    int npre = 0, nlanes = 1, nst = 1;
    long ld = 0, ldv = 0;
    int holeHiV = 0, ncT = 0;
    size_t matBytes = 0, wBytes = 0, lmatBytes = 0;
    int capPoints = 0, capLoocv = 0;
    int maxUnits = 0;

    Driver(fgp_ctx* parent_, fgp_ctx* wc_, Shared& S_, const std::vector<int>& order_) : parent(parent_), wc(wc_), S(S_), order(order_) {}

    // ------------------------------------------------------------------------------------------ set-up
    int setup() {
        DevState& dv = wc->dev[0];
        LCK(cudaSetDevice(dv.device));
        if (!wc->lockState) { wc->lockState = new LockState(); wc->lockFree = lock_free; }
        L = (LockState*)wc->lockState;
        n = S.n_rep > 0 ? (int)S.trRows[0].size() : S.n;
        npre = std::max(S.n_presample, S.n_starts);
        ld = fgp_pick_ld(n + 1);
        ncT = cdiv(n, FGP_TILE);
        matBytes = sizeof(double) * (size_t)ld * n;
        wBytes = sizeof(double) * (size_t)ncT * FGP_TILE * FGP_TILE;
        holeHiV = (int)fgp_roundup(n + 1, FGP_TILE);
        ldv = fgp_pick_ld(holeHiV + n);
        lmatBytes = sizeof(double) * (size_t)ldv * n;
        nlanes = wc->fit_lanes > 0 ? wc->fit_lanes : (n < 2048 ? 3 : 1);      // measured at n = 1024: 2 lanes 35.3, 3 lanes 35.8 candidates/s
        nst = n < 2048 ? 1 : (n < 4096 ? std::min(2, wc->streams) : wc->streams);
        if (nlanes > 1) nst = std::min(nst, DevState::MAXSTREAMS / nlanes);
        if (!L->streamsReady) {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            for (Lane& ln : L->lane) {
                for (int s = 0; s < DevState::MAXSTREAMS; ++s) {
                    LCK(cudaStreamCreateWithPriority(&ln.st[s], cudaStreamNonBlocking, hi));
                    LCK(cudaEventCreateWithFlags(&ln.evJoin[s], cudaEventDisableTiming));
                }
                LCK(cudaEventCreateWithFlags(&ln.evFork, cudaEventDisableTiming));
                LCK(cudaEventCreateWithFlags(&ln.evDone, cudaEventDisableTiming | (wc->blockingSync == 1 ? cudaEventBlockingSync : 0)));
            }
            L->streamsReady = true;
        }
        // ---- how many points fit: a share of what the device has left (workspaces of earlier calls are reused)
        int maxNeed = npre, dmaxUnits = S.ds;
        for (int j = 0; j < S.df; ++j) dmaxUnits += S.k[j];
        maxUnits = std::max(1, dmaxUnits);
        for (int c = 0; c < S.n_cand; ++c) {          // the largest finite-difference batch any structure of the call needs
            const int* ant = S.ants + (size_t)c * S.antLen;
            int dc = 0;
            for (int i = 0; i < S.ds; ++i) dc += ant[i] == 1;
            for (int j = 0; j < S.df; ++j) {
                const int* a = ant + S.ds + 4 * j;
                if (a[0] == 1) dc += a[1] == FGP_DIST_BYINDEX ? (a[2] > 0 ? a[2] : S.k[j]) : 1;
            }
            if (dc <= S.dmax) maxNeed = std::max(maxNeed, S.n_starts * (2 * dc + 1));
        }
        size_t freeB = 0, totB = 0;
        cudaMemGetInfo(&freeB, &totB);
        size_t have = 0;
        for (int l = 0; l < 4; ++l) have += L->lane[l].work.cap + L->lane[l].wbuf.cap + L->lane[l].lwork.cap + L->lane[l].lwbuf.cap;
        const size_t budget = std::min<size_t>((size_t)((freeB + have) * 0.6), (size_t)64 << 30) / nlanes;
        const size_t perPoint = matBytes + wBytes;
        const size_t perLoocv = lmatBytes + wBytes;
        // three quarters of the lane's budget for likelihood points, one quarter for leave-one-out trapezoids
        long pts = (long)(budget / 4 * 3 / perPoint), lv = (long)(budget / 4 / perLoocv);
        const int want = wc->fit_points > 0 ? wc->fit_points : 1024;
        capPoints = (int)std::max<long>(1, std::min<long>(pts, want));
        capLoocv = (int)std::max<long>(1, std::min<long>(lv, 64));
        if (capPoints < maxNeed) {
            if (pts < maxNeed) { fgp_set_err(wc, "fgp_fit_batch: not enough device memory for one candidate's likelihood batch"); return FGP_ERR_ALLOC; }
            capPoints = maxNeed;
        }
        for (int l = 0; l < nlanes; ++l) {
            Lane& ln = L->lane[l];
            if (ln.work.ensure(perPoint == 0 ? 1 : matBytes * capPoints) || ln.wbuf.ensure(wBytes * capPoints) ||
                ln.lwork.ensure(lmatBytes * capLoocv) || ln.lwbuf.ensure(wBytes * capLoocv)) {
                fgp_set_err(wc, "fgp_fit_batch: out of device memory (lock-step workspaces)");
                return FGP_ERR_ALLOC;
            }
            // staging of a full round up front: growing a device buffer later would synchronise the whole device
            const size_t blobMax = al16(sizeof(AsmElem) * (size_t)capPoints) + al16(sizeof(AsmElem) * (size_t)capLoocv) +
                                   al16(sizeof(AsmGroup) * (size_t)capPoints) +
                                   al16(sizeof(FgpUnit) * (size_t)maxUnits * (size_t)(capPoints + capLoocv)) +
                                   sizeof(double) * (size_t)(capPoints + capLoocv) * std::max(1, S.dmax) + 64;
            const size_t resMax = (sizeof(double) * 2 + sizeof(int)) * (size_t)capPoints + (sizeof(double) * 2 * (size_t)n + sizeof(int)) * capLoocv + 64;
            if (ln.ensure_pinned(&ln.hBlob, &ln.hBlobCap, blobMax) || ln.blob.ensure(blobMax) ||
                ln.ensure_pinned(&ln.hRes, &ln.hResCap, resMax) || ln.res.ensure(resMax)) {
                fgp_set_err(wc, "fgp_fit_batch: out of memory (round staging)");
                return FGP_ERR_ALLOC;
            }
        }
        return setup_arena();
    }

    // coordinate arena: every distinct column block the call's structures can use, sized up front
    int setup_arena() {
        Arena& A = L->arena;
        const int nrep = std::max(1, S.n_rep);
        A.col0.clear(); A.used = 0; A.rows = n;
        std::map<std::tuple<int, int, int, int>, int> widths;     // (input, p, basis, dist) -> columns
        for (int c = 0; c < S.n_cand; ++c) {
            const int* ant = S.ants + (size_t)c * S.antLen;
            for (int j = 0; j < S.df; ++j) {
                const int* a = ant + S.ds + 4 * j;
                if (a[0] != 1 || a[2] < 0 || a[2] > S.k[j]) continue;
                widths[std::make_tuple(j, a[2], a[3] == FGP_BASIS_PCA ? FGP_BASIS_PCA : FGP_BASIS_BSPLINES, a[1])] = a[2] ? a[2] : S.k[j];
            }
        }
        long cols = S.ds;
        for (auto& w : widths) cols += w.second;
        A.capCols = std::max<long>(1, cols * nrep);
        if (A.buf.ensure(sizeof(double) * (size_t)A.rows * A.capCols)) { fgp_set_err(wc, "fgp_fit_batch: out of device memory (coordinates)"); return FGP_ERR_ALLOC; }
        if (A.tmp.ensure(sizeof(FgpUnit) * (size_t)maxUnits + 64)) return FGP_ERR_ALLOC;
        // output vectors: the full y, or one per hold-out replicate
        if (A.ybuf.ensure(sizeof(double) * (size_t)n * nrep)) return FGP_ERR_ALLOC;
        A.dY.assign(nrep, nullptr);
        cudaStream_t up = wc->dev[0].sBulk;
        for (int r = 0; r < nrep; ++r) {
            A.dY[r] = (double*)A.ybuf.p + (size_t)r * n;
            const double* src = S.n_rep > 0 ? S.yTr[r].data() : S.y;
            LCK(cudaMemcpyAsync(A.dY[r], src, sizeof(double) * n, cudaMemcpyHostToDevice, up));
        }
        LCK(cudaStreamSynchronize(up));
        return 0;
    }

    // first arena column of a block, uploading it on first use (synchronous: a few dozen distinct blocks per call)
    int arena_block(const std::tuple<int, int, int, int, int>& key, const double* hostCols, int ncols, long* col) {
        Arena& A = L->arena;
        auto it = A.col0.find(key);
        if (it != A.col0.end()) { *col = it->second; return 0; }
        if (A.used + ncols > A.capCols) { fgp_set_err(wc, "fgp_fit_batch: coordinate arena overflow"); return FGP_ERR_STATE; }
        double* dst = (double*)A.buf.p + (size_t)A.used * A.rows;
        cudaStream_t up = wc->dev[0].sBulk;
        LCK(cudaMemcpyAsync(dst, hostCols, sizeof(double) * (size_t)A.rows * ncols, cudaMemcpyHostToDevice, up));
        LCK(cudaStreamSynchronize(up));
        *col = A.used;
        A.col0[key] = A.used;
        A.used += ncols;
        return 0;
    }

    // ------------------------------------------------------------------------------------------ admission
    // decode, place the coordinates, bounds (setBounds_*) and presample points (setSPoints_*) of one work item
    int admit(Cand& c, int item) {
        c.item = item;
        c.D = decode_item(S, wc, item);
        if (c.D.code) { crash(S, item, c.D.code); return 1; }
        const Decoded& D = c.D;
        const int rep = D.rep;
        c.n = D.n; c.ker = D.ker; c.d = D.d;
        if (c.d > S.dmax) { crash(S, item, -4); return 1; }
        c.units.clear(); c.lo.assign(c.d, 1e-10); c.up.assign(c.d, 0.0);
        int term = 0;
        std::vector<double> colbuf;
        for (int i : D.sIdx) {
            long col = 0;
            const auto key = std::make_tuple(-1 - i, 0, 0, 0, rep);
            if (!L->arena.col0.count(key)) {
                colbuf.resize(c.n);
                if (rep < 0) memcpy(colbuf.data(), S.sIn + (size_t)i * S.n, sizeof(double) * c.n);
                else for (int r = 0; r < c.n; ++r) colbuf[r] = S.sIn[S.trRows[rep][r] + (size_t)i * S.n];
            }
            int rc = arena_block(key, colbuf.data(), 1, &col);
            if (rc) return rc;
            // max |x_i - x_j| = max - min of the column (abs(outer(-)), R/7_distanceFunctions.R:31)
            double lo = 0, hi = 0;
            for (int r = 0; r < c.n; ++r) {
                const double x = rep < 0 ? S.sIn[r + (size_t)i * S.n] : S.sIn[S.trRows[rep][r] + (size_t)i * S.n];
                if (r == 0) lo = hi = x;
                lo = std::min(lo, x); hi = std::max(hi, x);
            }
            c.up[term] = 2.0 * (hi - lo);
            c.units.push_back({0, (int)col, term++, 1});
        }
        for (int q = 0; q < D.dfA; ++q) {
            Proj* pr = const_cast<Proj*>(D.proj[q]);
            const int j = D.fIdx[q], pe = pr->pe, dist = D.dis[q];
            const int* a = S.ants + (size_t)D.cand * S.antLen + S.ds + 4 * j;
            const auto key = std::make_tuple(j, a[2], a[3] == FGP_BASIS_PCA ? FGP_BASIS_PCA : FGP_BASIS_BSPLINES, dist, rep);
            long col = 0;
            if (dist == FGP_DIST_BYINDEX) {
                int rc = arena_block(key, pr->coefs.data(), pe, &col);
                if (rc) return rc;
                for (int b = 0; b < pe; ++b) {
                    c.up[term] = 2.0 * pr->colRange[b];
                    c.units.push_back({0, (int)(col + b), term++, 1});
                }
            } else {
                {   // whitened coefficients, computed once per projection (same arithmetic as build_coords)
                    std::lock_guard<std::mutex> lk(S.mu);
                    if (pr->z.empty()) {
                        whitening_factor(pr->J.data(), pe, pr->W);
                        pr->z.assign((size_t)c.n * pe, 0.0);
                        whiten_columns(pr->coefs.data(), c.n, pe, pr->W.data(), pr->z.data());
                    }
                }
                int rc = arena_block(key, pr->z.data(), pe, &col);
                if (rc) return rc;
                bool needMax;
                { std::lock_guard<std::mutex> lk(S.mu); needMax = pr->groupMax < 0.0; }
                if (needMax) {
                    // max group distance over the training curves: the kernel fgp_dist_max runs (setBounds_F)
                    std::vector<FgpUnit> tu;
                    for (int b = 0; b < pe; ++b) tu.push_back({1, (int)(col + b), 0, b == pe - 1});
                    cudaStream_t up = wc->dev[0].sBulk;
                    double* dOut = (double*)((char*)L->arena.tmp.p + sizeof(FgpUnit) * (size_t)maxUnits);
                    LCK(cudaMemcpyAsync(L->arena.tmp.p, tu.data(), sizeof(FgpUnit) * tu.size(), cudaMemcpyHostToDevice, up));
                    LCK(cudaMemsetAsync(dOut, 0, sizeof(double), up));
                    launch_distmax((const double*)L->arena.buf.p, c.n, c.n, (const FgpUnit*)L->arena.tmp.p, 0, pe, dOut, up);
                    double h = 0.0;
                    LCK(cudaMemcpyAsync(&h, dOut, sizeof(double), cudaMemcpyDeviceToHost, up));
                    LCK(cudaStreamSynchronize(up));
                    std::lock_guard<std::mutex> lk(S.mu);
                    pr->groupMax = h;
                }
                double gm;
                { std::lock_guard<std::mutex> lk(S.mu); gm = pr->groupMax; }
                c.up[term] = 2.0 * gm;
                for (int b = 0; b < pe; ++b) c.units.push_back({1, (int)(col + b), term, b == pe - 1});
                ++term;
            }
        }
        // presample points from the caller's uniform draws (R's runif, in ant order)
        const double* u = S.u01 + S.u01_off[item];
        c.pts.resize((size_t)c.d * npre);
        for (int j = 0; j < npre; ++j)
            for (int i = 0; i < c.d; ++i) c.pts[i + (size_t)j * c.d] = c.lo[i] + u[i + (size_t)j * c.d] * (c.up[i] - c.lo[i]);
        c.phase = Cand::PRESAMPLE;
        c.need = std::max(npre, S.n_starts * (2 * c.d + 1));
        c.have = false;
        c.mach.clear();
        return 0;
    }

    // ------------------------------------------------------------------------------------------ one round
    void build_request(Cand& c) {
        const int d = c.d;
        c.lslot = -1;
        if (c.phase == Cand::PRESAMPLE) { c.req = c.pts; c.nreq = npre; return; }
        if (c.phase == Cand::FINAL) { c.req = c.bestx; c.nreq = 1; return; }
        if (c.phase == Cand::SCORE) { c.req.clear(); c.nreq = 0; return; }
        // finite-difference points of every running start: x, then x +- 1e-3 e_i clipped at the box (optim's fmingr)
        const int nfd = 2 * d + 1;
        c.req.clear();
        for (auto& m : c.mach) {
            if (m.done) continue;
            const size_t base = c.req.size();
            c.req.resize(base + (size_t)d * nfd);
            double* fdp = c.req.data() + base;
            for (int j = 0; j < nfd; ++j)
                for (int i = 0; i < d; ++i) fdp[i + (size_t)j * d] = m.x[i];
            for (int i = 0; i < d; ++i) {
                fdp[i + (size_t)(1 + 2 * i) * d] = std::min(m.x[i] + 1e-3, c.up[i]);
                fdp[i + (size_t)(2 + 2 * i) * d] = std::max(m.x[i] - 1e-3, c.lo[i]);
            }
        }
        c.nreq = (int)(c.req.size() / d);
    }

    void finish(Cand& c, double q2) {
        const int it = c.item;
        S.fitness[it] = q2; S.nll[it] = c.bestf; S.sig2[it] = c.sig2; S.info[it] = 0;
        for (int i = 0; i < S.dmax; ++i) S.theta[(size_t)it * S.dmax + i] = i < c.d ? c.bestx[i] : std::nan("");
    }

    // returns true when the candidate has left the pipeline (scored or crashed)
    bool consume(Cand& c, const double* nll, const double* s2, const int* info) {
        const int d = c.d;
        if (c.phase == Cand::PRESAMPLE) {
            for (int j = 0; j < npre; ++j)
                if (info[j]) { crash(S, c.item, info[j]); return true; }      // quirk Q6: presampling does not catch chol errors
            std::vector<int> ord(npre);
            for (int j = 0; j < npre; ++j) ord[j] = j;
            std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return nll[a] < nll[b]; });
            c.mach.resize(S.n_starts);
            for (int s = 0; s < S.n_starts; ++s) {
                std::vector<double> x(c.pts.begin() + (size_t)ord[s] * d, c.pts.begin() + (size_t)(ord[s] + 1) * d);
                c.mach[s].init(x, c.lo, c.up);
            }
            c.phase = Cand::OPT;
            return false;
        }
        if (c.phase == Cand::OPT) {
            const int nfd = 2 * d + 1;
            int blk = 0;
            std::vector<double> g(d);
            for (auto& m : c.mach) {
                if (m.done) continue;
                const double* fv = nll + (size_t)blk * nfd;
                const int* fi = info + (size_t)blk * nfd;
                const double* fdp = c.req.data() + (size_t)blk * nfd * d;
                ++blk;
                bool ok = true;
                for (int j = 0; j < nfd; ++j)
                    if (fi[j] || !std::isfinite(fv[j])) ok = false;
                if (ok)
                    for (int i = 0; i < d; ++i)
                        g[i] = (fv[1 + 2 * i] - fv[2 + 2 * i]) / (fdp[i + (size_t)(1 + 2 * i) * d] - fdp[i + (size_t)(2 + 2 * i) * d]);
                m.feed(ok, fv[0], g);
            }
            for (auto& m : c.mach)
                if (!m.done) return false;
            // all starts have stopped: the best of those that did not hit an error (tryCatch per start,
            // R/3_training_SF.R:141-153; with one start the error reaches fitNtests_ACO and the ant is a crash)
            for (auto& m : c.mach) {
                if (m.res.failed_eval) continue;
                if (!c.have || m.res.f < c.bestf) { c.have = true; c.bestf = m.res.f; c.bestx = m.x; }
            }
            if (!c.have) { crash(S, c.item, -5); return true; }                // "All model optimizations crashed!"
            c.phase = Cand::FINAL;
            c.need = 1;
            return false;
        }
        if (c.phase == Cand::FINAL) {
            if (info[0]) { crash(S, c.item, info[0]); return true; }
            c.sig2 = s2[0];                                                    // variance at the optimum (:215-220)
            c.phase = Cand::SCORE;
            c.need = 0;
            return false;
        }
        return false;
    }

    // hold-out score of one (candidate, replicate): preMats on the training split, predict the validation rows
    // (R/1_Xfgpm_Class.R:545-549) -- through the single-model entry points, once per item
    int score_holdout(Cand& c) {
        const Decoded& D = c.D;
        const int rep = D.rep;
        std::vector<double> sAct, sVl;
        for (int i : D.sIdx) {
            for (int r : S.trRows[rep]) sAct.push_back(S.sIn[r + (size_t)i * S.n]);
            for (int r : S.vlRows[rep]) sVl.push_back(S.sIn[r + (size_t)i * S.n]);
        }
        std::vector<const double*> coefs, Js, coefsVl;
        for (const Proj* pr : D.proj) { coefs.push_back(pr->coefs.data()); Js.push_back(pr->J.data()); coefsVl.push_back(pr->coefsVl.data()); }
        fgp_problem* P = nullptr;
        int rc = fgp_problem_create(wc, c.n, D.dsA, D.dsA ? sAct.data() : nullptr, D.dfA, D.p.data(), D.dis.data(), coefs.data(), Js.data(),
                                    S.yTr[rep].data(), &P);
        if (rc) return rc;
        rc = fgp_premats(P, c.ker, S.nugget, c.sig2, c.bestx.data(), nullptr, nullptr);
        const int nvl = (int)S.vlRows[rep].size();
        std::vector<double> mean(nvl), sd(nvl);
        if (!rc) rc = fgp_predict(P, nvl, D.dsA ? sVl.data() : nullptr, coefsVl.data(), FGP_DETAIL_LIGHT, mean.data(), sd.data(), nullptr, nullptr);
        fgp_problem_destroy(P);
        if (rc > 0) { crash(S, c.item, rc); return 0; }
        if (rc < 0) return rc;
        finish(c, q2_of(S.yVl[rep].data(), mean.data(), nvl));
        return 0;
    }

    static size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

    // gather the lane's requests into one blob, launch the round on the lane's streams
    int launch_round(Lane& ln) {
        // ---- slots, grouped by correlation family (the assembly kernel is specialised on it); inside a family first the
        // candidates whose points can share one batch-fused assembly (a group = the consecutive points of one candidate:
        // same coordinates, different length-scales), then the rest
        int B = 0, Lc = 0;
        int kerEnd[4] = {0, 0, 0, 0}, lkerEnd[4] = {0, 0, 0, 0}, fusedEnd[4] = {0, 0, 0, 0}, groupEnd[4] = {0, 0, 0, 0};
        size_t nScale = 0, nUnits = 0;
        std::vector<AsmGroup> groups;
        const bool useFused = wc->fused_assembly != 0;
        auto fused = [useFused](const Cand* c) { return useFused && c->nreq >= 4 && assemble_fusable(c->d, (int)c->units.size()); };
        for (int ker = 1; ker <= 3; ++ker) {
            for (Cand* c : ln.cands)
                if (c->ker == ker && c->nreq > 0 && fused(c)) {
                    c->slot0 = B; B += c->nreq; nScale += (size_t)c->nreq * c->d;
                    groups.push_back({c->slot0, c->nreq});
                }
            fusedEnd[ker] = B; groupEnd[ker] = (int)groups.size();
            for (Cand* c : ln.cands)
                if (c->ker == ker && c->nreq > 0 && !fused(c)) { c->slot0 = B; B += c->nreq; nScale += (size_t)c->nreq * c->d; }
            kerEnd[ker] = B;
        }
        for (int ker = 1; ker <= 3; ++ker) {
            for (Cand* c : ln.cands)
                if (c->ker == ker && c->phase == Cand::SCORE && S.n_rep == 0 && Lc < capLoocv) { c->lslot = Lc++; nScale += c->d; }
            lkerEnd[ker] = Lc;
        }
        for (Cand* c : ln.cands) nUnits += c->units.size();
        ln.Btot = B; ln.Lcnt = Lc;
        if (B == 0 && Lc == 0) return 0;
        FGP_COUNT_WORK(0, B); FGP_COUNT_WORK(1, Lc);
        // ---- blob: elements | loocv elements | fused groups | unit tables | length-scale multipliers
        const size_t offElem = 0, offLelem = al16(sizeof(AsmElem) * (size_t)B), offGroups = offLelem + al16(sizeof(AsmElem) * (size_t)Lc);
        const size_t offUnits = offGroups + al16(sizeof(AsmGroup) * groups.size());
        const size_t offScale = offUnits + al16(sizeof(FgpUnit) * nUnits), blobBytes = offScale + sizeof(double) * nScale;
        if (ln.ensure_pinned(&ln.hBlob, &ln.hBlobCap, blobBytes) || ln.blob.ensure(std::max<size_t>(blobBytes, 1 << 16))) {
            fgp_set_err(wc, "fgp_fit_batch: out of memory (round blob)");
            return FGP_ERR_ALLOC;
        }
        char* dB = (char*)ln.blob.p;
        AsmElem* hE = (AsmElem*)(ln.hBlob + offElem);
        AsmElem* hLE = (AsmElem*)(ln.hBlob + offLelem);
        FgpUnit* hU = (FgpUnit*)(ln.hBlob + offUnits);
        double* hS = (double*)(ln.hBlob + offScale);
        if (!groups.empty()) memcpy(ln.hBlob + offGroups, groups.data(), sizeof(AsmGroup) * groups.size());
        const AsmGroup* dG = (const AsmGroup*)(dB + offGroups);
        const double* arenaX = (const double*)L->arena.buf.p;
        size_t uPos = 0, sPos = 0;
        for (Cand* c : ln.cands) {
            if (c->nreq == 0 && c->lslot < 0) continue;
            const FgpUnit* dU = (const FgpUnit*)(dB + offUnits) + uPos;
            memcpy(hU + uPos, c->units.data(), sizeof(FgpUnit) * c->units.size());
            uPos += c->units.size();
            const double* dy = L->arena.dY[c->D.rep < 0 ? 0 : c->D.rep];
            for (int j = 0; j < c->nreq; ++j) {
                kernel_scales(c->ker, c->d, c->req.data() + (size_t)j * c->d, hS + sPos);
                hE[c->slot0 + j] = AsmElem{arenaX, dU, (const double*)(dB + offScale) + sPos, dy, (int)c->units.size(), 0};
                sPos += c->d;
            }
            if (c->lslot >= 0) {
                kernel_scales(c->ker, c->d, c->bestx.data(), hS + sPos);
                hLE[c->lslot] = AsmElem{arenaX, dU, (const double*)(dB + offScale) + sPos, dy, (int)c->units.size(), 0};
                sPos += c->d;
            }
        }
        // ---- results: nll | sig2 | info  and  diag | Ry | info of the leave-one-out batch
        const size_t offSig = sizeof(double) * (size_t)B, offInfo = 2 * offSig, offDiag = al16(offInfo + sizeof(int) * (size_t)B);
        const size_t offRy = offDiag + sizeof(double) * (size_t)Lc * n, offLinfo = offRy + sizeof(double) * (size_t)Lc * n;
        const size_t resBytes = offLinfo + sizeof(int) * (size_t)Lc;
        if (ln.ensure_pinned(&ln.hRes, &ln.hResCap, resBytes) || ln.res.ensure(std::max<size_t>(resBytes, 1 << 16))) {
            fgp_set_err(wc, "fgp_fit_batch: out of memory (round results)");
            return FGP_ERR_ALLOC;
        }
        char* dR = (char*)ln.res.p;
        double* dNll = (double*)dR; double* dSig = (double*)(dR + offSig); int* dInfo = (int*)(dR + offInfo);
        double* dDiag = (double*)(dR + offDiag); double* dRy = (double*)(dR + offRy); int* dLinfo = (int*)(dR + offLinfo);
        const AsmElem* dE = (const AsmElem*)(dB + offElem);
        const AsmElem* dLE = (const AsmElem*)(dB + offLelem);

        cudaStream_t s0 = ln.st[0];
        LCK(cudaMemcpyAsync(dB, ln.hBlob, blobBytes, cudaMemcpyHostToDevice, s0));
        LCK(cudaMemsetAsync(dInfo, 0, resBytes - offInfo, s0));
        // ---- assembly of every point of the round on the lane's first stream: one batch-fused launch per correlation
        // family over the candidates' groups, one plain launch for the rest
        for (int ker = 1; ker <= 3 && B > 0; ++ker) {
            AsmArgs a{};
            a.XP = arenaX; a.nP = n; a.ldP = n; a.XQ = arenaX; a.nQ = n; a.ldQ = n;
            a.ker = ker; a.sym = 1; a.mult = 1.0; a.diag_rel = S.nugget; a.diag_abs = 0.0;
            a.ld = ld; a.strideM = (long)ld * n;
            FgpTimers* tm = wc->profile ? &wc->dev[0].timers : nullptr;
            if (groupEnd[ker] > groupEnd[ker - 1]) {
                a.M = (double*)ln.work.p; a.elems = dE;                  // groups carry absolute slot numbers
                if (tm) tm->begin(s0);
                launch_assemble_fused(a, dG + groupEnd[ker - 1], groupEnd[ker] - groupEnd[ker - 1], 0, s0);
                if (tm) tm->end(s0, FgpTimers::ASSEMBLY, 0.0);
            }
            if (kerEnd[ker] > fusedEnd[ker]) {
                a.M = (double*)ln.work.p + (size_t)fusedEnd[ker] * ld * n; a.elems = dE + fusedEnd[ker];
                if (tm) tm->begin(s0);
                launch_assemble(a, kerEnd[ker] - fusedEnd[ker], s0);
                if (tm) tm->end(s0, FgpTimers::ASSEMBLY, 0.0);
            }
        }
        // ---- factorisations: groups of slots on the lane's streams (one group below n = 2048)
        // (dataflow Cholesky: all slots of the round in one persistent launch on the lane's first stream)
        const bool useDag = B > 0 && wc->dag && !wc->profile && chol_dag_supported(n, n + 1) && !ln.dag.ensure(chol_dag_flag_bytes(n, B));
        const int ns = B > 0 ? (useDag ? 1 : std::max(1, std::min(nst, B))) : 0;
        if (ns > 1) LCK(cudaEventRecord(ln.evFork, s0));
        const int per = ns ? cdiv(B, ns) : 0;
        for (int s = 0; s < ns; ++s) {
            const int g0 = s * per, g1 = std::min(B, g0 + per);
            if (g1 <= g0) continue;
            cudaStream_t st = ln.st[s];
            if (s > 0) LCK(cudaStreamWaitEvent(st, ln.evFork, 0));
            double* Mg = (double*)ln.work.p + (size_t)g0 * ld * n;
            launch_fill_yrow_elems(Mg, ld, (long)ld * n, g1 - g0, n, dE + g0, n, st);
            TrapDesc t{};
            t.M = Mg; t.ld = ld; t.strideM = (long)ld * n; t.batch = g1 - g0;
            t.ncols = n; t.nrows = n + 1; t.holeLo = n + 1; t.holeHi = n + 1; t.colHoleLo = n; t.colHoleHi = n; t.cpre = 0; t.upperRows = 0;
            t.W = (double*)ln.wbuf.p + (size_t)g0 * ncT * FGP_TILE * FGP_TILE; t.strideW = (long)ncT * FGP_TILE * FGP_TILE;
            t.info = dInfo + g0;
            if (!(useDag && launch_chol_dag(t.M, t.ld, t.strideM, t.batch, n, t.W, t.strideW, t.info, ln.dag.p, st, 1, wc->dag_order))) {
                int rc = trap_factor(wc, 0, t, st, nullptr, false);
                if (rc) return rc;
            }
            launch_nll_reduce(Mg, ld, (long)ld * n, g1 - g0, n, n, nullptr, dInfo + g0, dNll + g0, dSig + g0, st);
            if (s > 0) {
                LCK(cudaEventRecord(ln.evJoin[s], st));
                LCK(cudaStreamWaitEvent(s0, ln.evJoin[s], 0));
            }
        }
        // ---- leave-one-out batch (getOut_loocv, R/8_outilsStats.R:1-7): R + 2 nugget I (quirk Q2) with identity rows
        if (Lc > 0) {
            double* Mv = (double*)ln.lwork.p;
            for (int ker = 1; ker <= 3; ++ker) {
                const int a0 = lkerEnd[ker - 1], a1 = lkerEnd[ker];
                if (a1 <= a0) continue;
                AsmArgs a{};
                a.XP = arenaX; a.nP = n; a.ldP = n; a.XQ = arenaX; a.nQ = n; a.ldQ = n;
                a.ker = ker; a.sym = 1; a.mult = 1.0; a.diag_rel = 2.0 * S.nugget; a.diag_abs = 0.0;
                a.M = Mv + (size_t)a0 * ldv * n; a.ld = ldv; a.strideM = (long)ldv * n;
                a.elems = dLE + a0;
                launch_assemble(a, a1 - a0, s0);
            }
            launch_fill_yrow_elems(Mv, ldv, (long)ldv * n, Lc, n, dLE, n, s0);
            launch_fill_identity_rows(Mv, ldv, (long)ldv * n, Lc, holeHiV, n, s0);
            TrapDesc t{};
            t.M = Mv; t.ld = ldv; t.strideM = (long)ldv * n; t.batch = Lc;
            t.ncols = n; t.nrows = holeHiV + n; t.holeLo = n + 1; t.holeHi = holeHiV; t.colHoleLo = n; t.colHoleHi = n;
            t.cpre = 0; t.upperRows = 1; t.W = (double*)ln.lwbuf.p; t.strideW = (long)ncT * FGP_TILE * FGP_TILE; t.info = dLinfo;
            int rc = trap_factor(wc, 0, t, s0, nullptr, false);
            if (rc) return rc;
            // X = L^-T in the extra rows: diag(Rinv)_i = |X_i.|^2 ; (Rinv y)_i = X_i. (L^-1 y)
            launch_rowstats(Mv, ldv, (long)ldv * n, Lc, holeHiV, n, n, 1, Mv + n, ldv, (long)ldv * n, dDiag, dRy, s0);
        }
        LCK(cudaMemcpyAsync(ln.hRes, dR, resBytes, cudaMemcpyDeviceToHost, s0));
        LCK(cudaEventRecord(ln.evDone, s0));
        LCK(cudaGetLastError());
        ln.busy = true;
        return 0;
    }

    int wait_lane(Lane& ln) {
        if (wc->blockingSync == 2) {
            for (;;) {
                cudaError_t e = cudaEventQuery(ln.evDone);
                if (e == cudaSuccess) break;
                if (e != cudaErrorNotReady) { fgp_set_err_cuda(wc, "cudaEventQuery", e, __FILE__, __LINE__); return FGP_ERR_CUDA; }
                sched_yield();
            }
        } else LCK(cudaEventSynchronize(ln.evDone));
        ln.busy = false;
        return 0;
    }

    // hand the round's results to the candidates; finished ones leave the lane
    int consume_lane(Lane& ln) {
        const int B = ln.Btot, Lc = ln.Lcnt;
        const size_t offSig = sizeof(double) * (size_t)B, offInfo = 2 * offSig, offDiag = al16(offInfo + sizeof(int) * (size_t)B);
        const size_t offRy = offDiag + sizeof(double) * (size_t)Lc * n, offLinfo = offRy + sizeof(double) * (size_t)Lc * n;
        const double* hNll = (const double*)ln.hRes; const double* hSig = (const double*)(ln.hRes + offSig);
        const int* hInfo = (const int*)(ln.hRes + offInfo);
        const double* hDiag = (const double*)(ln.hRes + offDiag); const double* hRy = (const double*)(ln.hRes + offRy);
        const int* hLinfo = (const int*)(ln.hRes + offLinfo);
        std::vector<Cand*> keep;
        std::vector<double> yh(n);
        for (Cand* c : ln.cands) {
            bool gone = false;
            if (c->lslot >= 0) {
                // y_pre = y - Rinv y / diag(Rinv) (R/8_outilsStats.R:6); Q2 (R/1_Xfgpm_Class.R:541)
                if (hLinfo[c->lslot]) crash(S, c->item, hLinfo[c->lslot]);
                else {
                    const double* dg = hDiag + (size_t)c->lslot * n; const double* ry = hRy + (size_t)c->lslot * n;
                    for (int i = 0; i < n; ++i) yh[i] = S.y[i] - ry[i] / dg[i];
                    finish(*c, q2_of(S.y, yh.data(), n));
                }
                gone = true;
            } else if (c->nreq > 0) {
                gone = consume(*c, hNll + c->slot0, hSig + c->slot0, hInfo + c->slot0);
                if (!gone && c->phase == Cand::SCORE && S.n_rep > 0) {
                    int rc = score_holdout(*c);
                    if (rc) return rc;
                    gone = true;
                }
            }
            if (gone) { FGP_COUNT_WORK(3, 1); delete c; } else keep.push_back(c);
        }
        ln.cands.swap(keep);
        return 0;
    }

    int run() {
        int rc = setup();
        if (rc) return rc;
        const int n_items = (int)order.size();
        Cand* pending = nullptr;          // admitted but not yet placed (did not fit the lane that asked)
        for (;;) {
            bool active = false;
            for (int l = 0; l < nlanes; ++l) {
                Lane& ln = L->lane[l];
                if (ln.busy) {
                    if ((rc = wait_lane(ln))) return rc;
                    if ((rc = consume_lane(ln))) return rc;
                }
                // top the lane up from the queue
                int needPts = 0;
                for (Cand* c : ln.cands) needPts += c->need;
                for (;;) {
                    if (!pending) {
                        const int q = S.next.fetch_add(1);
                        if (q >= n_items) break;
                        Cand* c = new Cand();
                        rc = admit(*c, order[q]);
                        if (rc < 0) { delete c; return rc; }
                        if (rc > 0) { delete c; continue; }       // crashed at decoding: outputs written
                        pending = c;
                    }
                    if (!ln.cands.empty() && needPts + pending->need > capPoints) break;
                    needPts += pending->need;
                    ln.cands.push_back(pending);
                    pending = nullptr;
                }
                if (!ln.cands.empty()) {
                    for (Cand* c : ln.cands) build_request(*c);
                    if ((rc = launch_round(ln))) return rc;
                    active = true;
                }
            }
            if (!active && !pending) break;
        }
        return 0;
    }
};

}  // namespace

int fit_lockstep(fgp_ctx* ctx, Shared& S, const std::vector<int>& order) {
    const int nd = (int)ctx->dev.size();
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (ctx->fitWorkers.size() < (size_t)nd * 16) ctx->fitWorkers.resize((size_t)nd * 16, nullptr);
    }
    std::vector<int> rcs(nd, 0);
    std::vector<std::thread> th;
    for (int dI = 0; dI < nd; ++dI) {
        th.emplace_back([&, dI]() {
            const int device = ctx->dev[dI].device;
            fgp_ctx** slot = &ctx->fitWorkers[(size_t)dI * 16];
            if (!*slot && fgp_ctx_create(1, &device, slot) != 0) { *slot = nullptr; rcs[dI] = FGP_ERR_CUDA; return; }
            fgp_ctx* wc = *slot;
            cudaSetDevice(device);
            wc->streams = ctx->streams; wc->outer = ctx->outer; wc->outer_single = ctx->outer_single; wc->lookahead = ctx->lookahead;
            wc->fit_points = ctx->fit_points; wc->fit_lanes = ctx->fit_lanes; wc->predict_reserve = ctx->predict_reserve;
            wc->fused_assembly = ctx->fused_assembly; wc->pdl = ctx->pdl; wc->gemm_tma = ctx->gemm_tma; wc->gemm_flat = ctx->gemm_flat; wc->dag = ctx->dag; wc->dag_order = ctx->dag_order;
            wc->blockingSync = ctx->fit_blocking > 0 ? ctx->fit_blocking : 0;
            // option "profile": per-launch CUDA events of this call's rounds on the first GPU (meaningful with fit_lanes = 1,
            // where the rounds do not overlap); fgp_get_timing on the caller's context reports them
            wc->profile = (ctx->profile && dI == 0) ? 1 : 0;
            if (wc->profile) wc->dev[0].timers.reset();
            Driver drv(ctx, wc, S, order);
            rcs[dI] = drv.run();
            if (wc->profile) { cudaDeviceSynchronize(); ctx->dev[0].timers.swap(wc->dev[0].timers); wc->profile = 0; }
            if (rcs[dI]) {
                // leave no stream work behind and no item unreported
                cudaDeviceSynchronize();
                S.infra(rcs[dI], wc);
                if (wc->lockState)
                    for (Lane& ln : ((LockState*)wc->lockState)->lane) {
                        for (Cand* c : ln.cands) { crash(S, c->item, -3); delete c; }
                        ln.cands.clear(); ln.busy = false;
                    }
            }
        });
    }
    for (auto& t : th) t.join();
    for (int dI = 0; dI < nd; ++dI)
        if (rcs[dI]) { fgp_set_err(ctx, "fgp_fit_batch (lock-step): " + S.infraMsg); return rcs[dI]; }
    return 0;
}
