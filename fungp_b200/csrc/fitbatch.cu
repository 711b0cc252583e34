// fgp_fit_batch: batched scoring of candidate model structures for the ant-colony search.
// One candidate = what fitNtests_ACO does per ant (R/3_ant_admin.R:810-831): formatSol_ACO decode (:301-373) ->
// fgpm(): dimReduction, distances, setBounds / setSPoints / optimHypers (R/3_training_SF.R:57-223), preMats
// (R/4_prediction_SF.R:24-32) -> getFitness: Q2 of the leave-one-out predictions (R/1_Xfgpm_Class.R:522-543).
// Host orchestration only: every number comes from the same C-ABI entry points a single fgpm() uses
// (fgp_project, fgp_problem_create, fgp_dist_max, fgp_nll_batch, fgp_premats, fgp_loocv); the optimiser is the host
// L-BFGS-B of lbfgsb.hpp with one batched device call per iterate (value + 2d finite-difference neighbours).
// Candidates are independent: they are dealt from a queue to worker threads, `fit_workers` per GPU of the context,
// each with its own streams and workspaces, so small-n fits overlap on the chip and GPUs never exchange data.
#include "fitcommon.h"
#include "lbfgsb.hpp"
#include <cstdlib>
#include <cstring>
#include <thread>

using namespace fgp_fit;

namespace {

void fit_one(Shared& S, fgp_ctx* ctx, int item) {
    const int c = item;                 // output slot
    // ---- formatSol_ACO
    const Decoded D = decode_item(S, ctx, item);
    if (D.code) { crash(S, c, D.code); return; }
    const int rep = D.rep, n = D.n, ker = D.ker, dsA = D.dsA, dfA = D.dfA;
    const double* yv = rep < 0 ? S.y : S.yTr[rep].data();
    std::vector<double> sAct, sVl;
    for (int i : D.sIdx) {
        if (rep < 0) sAct.insert(sAct.end(), S.sIn + (size_t)i * S.n, S.sIn + (size_t)(i + 1) * S.n);
        else {
            for (int r : S.trRows[rep]) sAct.push_back(S.sIn[r + (size_t)i * S.n]);
            for (int r : S.vlRows[rep]) sVl.push_back(S.sIn[r + (size_t)i * S.n]);
        }
    }
    std::vector<const double*> coefs, Js, coefsVl;
    for (const Proj* pr : D.proj) { coefs.push_back(pr->coefs.data()); Js.push_back(pr->J.data()); coefsVl.push_back(pr->coefsVl.data()); }
    const std::vector<int>& p = D.p;
    const std::vector<int>& dis = D.dis;

    fgp_problem* P = nullptr;
    int rc = fgp_problem_create(ctx, n, dsA, dsA ? sAct.data() : nullptr, dfA, p.data(), dis.data(), coefs.data(), Js.data(), yv, &P);
    if (rc) { S.infra(rc, ctx); crash(S, c, -3); return; }
    const int d = fgp_problem_nls(P);
    auto done = [&](int code) {
        if (code == -3) S.infra(rc < 0 ? rc : FGP_ERR_CUDA, ctx);     // the library itself failed: not a model crash
        if (code) crash(S, c, code);
        fgp_problem_destroy(P);
    };
    if (d > S.dmax) { done(-4); return; }
    // ---- setBounds_*: lower 1e-10, upper 2 max(D_l)
    std::vector<double> lo(d, 1e-10), up(d);
    if ((rc = fgp_dist_max(P, up.data()))) { done(-3); return; }
    for (int i = 0; i < d; ++i) up[i] *= 2.0;
    // ---- setSPoints_*: the presample draws were made by the caller (R's runif, in ant order)
    const int npre = std::max(S.n_presample, S.n_starts);
    const double* u = S.u01 + S.u01_off[item];
    std::vector<double> pts((size_t)d * npre), v(npre), s2(npre);
    std::vector<int> inf(npre);
    for (int j = 0; j < npre; ++j)
        for (int i = 0; i < d; ++i) pts[i + (size_t)j * d] = lo[i] + u[i + (size_t)j * d] * (up[i] - lo[i]);
    if ((rc = fgp_nll_batch(P, ker, S.nugget, npre, pts.data(), nullptr, v.data(), s2.data(), inf.data()))) { done(-3); return; }
    for (int j = 0; j < npre; ++j)
        if (inf[j]) { done(inf[j]); return; }       // quirk Q6: a chol failure while presampling aborts the fit
    std::vector<int> ord(npre);
    for (int j = 0; j < npre; ++j) ord[j] = j;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return v[a] < v[b]; });
    // ---- optimHypers_*: L-BFGS-B from each start, keep the best
    const int nfd = 2 * d + 1;
    std::vector<double> fdp((size_t)d * nfd), fv(nfd), fs(nfd);
    std::vector<int> fi(nfd);
    bool infra = false;
    fgp_lbfgsb::FG fg = [&](const std::vector<double>& x, double& f, std::vector<double>& g) {
        for (int j = 0; j < nfd; ++j)
            for (int i = 0; i < d; ++i) fdp[i + (size_t)j * d] = x[i];
        for (int i = 0; i < d; ++i) {
            fdp[i + (size_t)(1 + 2 * i) * d] = std::min(x[i] + 1e-3, up[i]);
            fdp[i + (size_t)(2 + 2 * i) * d] = std::max(x[i] - 1e-3, lo[i]);
        }
        if ((rc = fgp_nll_batch(P, ker, S.nugget, nfd, fdp.data(), nullptr, fv.data(), fs.data(), fi.data()))) { infra = true; return false; }
        for (int j = 0; j < nfd; ++j)
            if (fi[j] || !std::isfinite(fv[j])) return false;
        f = fv[0];
        g.resize(d);
        for (int i = 0; i < d; ++i)
            g[i] = (fv[1 + 2 * i] - fv[2 + 2 * i]) / (fdp[i + (size_t)(1 + 2 * i) * d] - fdp[i + (size_t)(2 + 2 * i) * d]);
        return true;
    };
    bool have = false;
    double bestf = 0.0;
    std::vector<double> bestx;
    for (int sI = 0; sI < S.n_starts; ++sI) {
        std::vector<double> x(pts.begin() + (size_t)ord[sI] * d, pts.begin() + (size_t)(ord[sI] + 1) * d);
        fgp_lbfgsb::Result r = fgp_lbfgsb::minimize(fg, x, lo, up);
        if (infra) { done(-3); return; }
        // Any error inside optim() aborts that start: the reference wraps every start in tryCatch when n.starts > 1
        // (R/3_training_SF.R:141-153) and, with a single start, lets the error reach fitNtests_ACO, which logs the ant as
        // a crash (R/3_ant_admin.R:815-829).  A not-positive-definite trial point or finite-difference neighbour therefore
        // never leaves a half-optimised iterate behind.
        if (r.failed_eval) continue;
        if (!have || r.f < bestf) { have = true; bestf = r.f; bestx = x; }
    }
    if (!have) { done(-5); return; }                        // "All model optimizations crashed!"
    // ---- variance at the optimum, preMats, Q2loocv
    double f1, s1; int i1 = 0;
    if ((rc = fgp_nll_batch(P, ker, S.nugget, 1, bestx.data(), nullptr, &f1, &s1, &i1)) || i1) { done(i1 ? i1 : -3); return; }
    double q2 = 0.0;
    if (rep < 0) {
        // Q2loocv (R/1_Xfgpm_Class.R:539-542).  getOut_loocv rebuilds R from the hyper-parameters, so the candidate's
        // preMats are only materialised for hold-out scoring (the caller rebuilds the winner's model from theta, sig2)
        rc = fgp_premats(P, ker, S.nugget, s1, bestx.data(), nullptr, nullptr);
        if (rc) { done(rc > 0 ? rc : -3); return; }
        rc = fgp_loocv(P, nullptr, &q2);
        if (rc) { done(rc > 0 ? rc : -3); return; }
    } else {
        rc = fgp_premats(P, ker, S.nugget, s1, bestx.data(), nullptr, nullptr);
        if (rc) { done(rc > 0 ? rc : -3); return; }
        // Q2hout: predict the validation rows (R/1_Xfgpm_Class.R:545-549)
        const int nvl = (int)S.vlRows[rep].size();
        std::vector<double> mean(nvl), sd(nvl);
        rc = fgp_predict(P, nvl, dsA ? sVl.data() : nullptr, coefsVl.data(), FGP_DETAIL_LIGHT, mean.data(), sd.data(), nullptr, nullptr);
        if (rc) { done(rc > 0 ? rc : -3); return; }
        q2 = q2_of(S.yVl[rep].data(), mean.data(), nvl);
    }
    S.fitness[c] = q2; S.nll[c] = bestf; S.sig2[c] = s1; S.info[c] = 0;
    for (int i = 0; i < S.dmax; ++i) S.theta[(size_t)c * S.dmax + i] = i < d ? bestx[i] : std::nan("");
    done(0);
}

}  // namespace

static int fit_batch_impl(fgp_ctx* ctx, int n, int ds, const double* sIn, int df, const int* k, const double* const* fIn,
                          const double* y, int n_cand, const int* ants, double nugget, int n_starts, int n_presample,
                          const double* u01, const long long* u01_off, int dmax, int n_rep, int n_vl, const int* ind_vl,
                          double* fitness, double* nll, double* sig2, double* theta, int* info) {
    if (!ctx || n <= 0 || ds < 0 || df < 0 || (ds > 0 && !sIn) || (df > 0 && (!k || !fIn)) || !y || n_cand < 0 || !ants ||
        n_starts < 1 || n_presample < 1 || !u01 || !u01_off || dmax < 1 || !fitness || !nll || !sig2 || !theta || !info)
        return FGP_ERR_ARG;
    if (n_rep < 0 || (n_rep > 0 && (n_vl < 1 || n_vl >= n || !ind_vl))) return FGP_ERR_ARG;
    Shared S;
    S.n = n; S.ds = ds; S.df = df; S.sIn = sIn; S.k = k; S.fIn = fIn; S.y = y;
    S.n_cand = n_cand; S.antLen = ds + 4 * df + 1; S.ants = ants;
    S.nugget = nugget; S.n_starts = n_starts; S.n_presample = n_presample;
    S.u01 = u01; S.u01_off = u01_off; S.dmax = dmax;
    S.fitness = fitness; S.nll = nll; S.sig2 = sig2; S.theta = theta; S.info = info;
    S.n_rep = n_rep; S.n_vl = n_vl; S.ind_vl = ind_vl;
    for (int r = 0; r < n_rep; ++r) {            // splitData (R/1_Xfgpm_Class.R:559-584)
        std::vector<char> isvl(n, 0);
        std::vector<int> vl, tr;
        for (int i = 0; i < n_vl; ++i) {
            const int idx = ind_vl[i + (size_t)r * n_vl] - 1;
            if (idx < 0 || idx >= n) return FGP_ERR_ARG;
            isvl[idx] = 1; vl.push_back(idx);
        }
        for (int i = 0; i < n; ++i) if (!isvl[i]) tr.push_back(i);
        std::vector<double> ytr, yvl;
        for (int i : tr) ytr.push_back(y[i]);
        for (int i : vl) yvl.push_back(y[i]);
        S.trRows.push_back(tr); S.vlRows.push_back(vl); S.yTr.push_back(ytr); S.yVl.push_back(yvl);
    }
    const int n_items = n_cand * std::max(1, n_rep);
    const int nd = (int)ctx->dev.size();
    const int wpd = std::max(1, ctx->fit_workers);
    // Longest first: the cost of a fit grows with its number of length-scales (finite-difference batch = 2d + 1 and
    // more L-BFGS-B iterations), so the queue hands out the big structures first and the small ones fill the tail.
    // Results are written per item, so the order of the outputs does not change.
    std::vector<int> order(n_items), nls(n_items, 0);
    for (int it = 0; it < n_items; ++it) {
        order[it] = it;
        const int* ant = ants + (size_t)(it / std::max(1, n_rep)) * S.antLen;
        int d = 0;
        for (int i = 0; i < ds; ++i) d += ant[i] == 1;
        for (int j = 0; j < df; ++j) {
            const int* a = ant + ds + 4 * j;
            if (a[0] != 1) continue;
            d += a[1] == FGP_DIST_BYINDEX ? (a[2] > 0 ? a[2] : k[j]) : 1;
        }
        nls[it] = d;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nls[a] > nls[b]; });
    if (ctx->fit_mode == 1) {
        const int rcl = fit_lockstep(ctx, S, order);
        cudaSetDevice(ctx->dev[0].device);
        if (rcl) return rcl;
        if (S.infraErr.load()) { fgp_set_err(ctx, "fgp_fit_batch: " + S.infraMsg); return S.infraErr.load(); }
        return 0;
    }
    std::vector<std::thread> th;
    std::atomic<int> ctxFail{0};
    // How the workers wait on the host: 0 spin (cudaStreamSynchronize), 1 sleep on a blocking-sync event, 2 poll the event
    // and sched_yield.  Measured, n = 1024 candidates: 1 GPU / 16 cores 29.1 (spin) 29.2 (yield) 26.6 (sleep) candidates/s;
    // 8 GPUs / 32 cores / 64 workers 147 (spin) 203-212 (yield) 196 (sleep).  Auto: yield when the workers of all visible
    // GPUs (one process per GPU under torchrun) outnumber 3/4 of the cores
    int blocking = ctx->fit_blocking;
    if (blocking < 0) {
        int visible = 1;
        if (cudaGetDeviceCount(&visible) != cudaSuccess) { cudaGetLastError(); visible = 1; }
        const unsigned cores = std::max(1u, std::thread::hardware_concurrency());
        // under torchrun LOCAL_WORLD_SIZE says how many such processes share the host (the GPUs may be hidden per rank)
        int peers = 0;
        if (const char* lws = getenv("LOCAL_WORLD_SIZE")) peers = atoi(lws);
        const int gpus = peers > 0 ? peers * nd : std::max(visible, nd);
        blocking = ((unsigned)(wpd * gpus) * 4u > cores * 3u) ? 2 : 0;
    }
    // worker contexts persist in the parent context (slot = device index * 16 + worker)
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (ctx->fitWorkers.size() < (size_t)nd * 16) ctx->fitWorkers.resize((size_t)nd * 16, nullptr);
    }
    for (int dI = 0; dI < nd; ++dI)
        for (int w = 0; w < wpd; ++w) {
            const int device = ctx->dev[dI].device;
            fgp_ctx** slot = &ctx->fitWorkers[(size_t)dI * 16 + w];
            th.emplace_back([&, device, slot, blocking]() {
                if (!*slot && fgp_ctx_create(1, &device, slot) != 0) { *slot = nullptr; ctxFail++; return; }
                fgp_ctx* wc = *slot;
                cudaSetDevice(device);
                wc->streams = ctx->streams; wc->outer = ctx->outer; wc->outer_single = ctx->outer_single; wc->lookahead = ctx->lookahead;
                wc->blockingSync = blocking; wc->graphs = ctx->graphs; wc->fused_assembly = ctx->fused_assembly; wc->pdl = ctx->pdl; wc->gemm_tma = ctx->gemm_tma; wc->gemm_flat = ctx->gemm_flat; wc->dag = ctx->dag; wc->dag_order = ctx->dag_order;
                for (;;) {
                    const int q = S.next.fetch_add(1);
                    if (q >= n_items) break;
                    fit_one(S, wc, order[q]);
                    FGP_COUNT_WORK(3, 1);
                }
                // large workspaces go back to the device (candidates at n = 8192 need GBs per worker); small ones stay
                DevState& d = wc->dev[0];
                if (d.work.cap + d.wbuf.cap > ((size_t)2 << 30)) { d.work.release(); d.wbuf.release(); }
            });
        }
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->dev[0].device);
    if (ctxFail.load() == nd * wpd && n_items > 0) { fgp_set_err(ctx, "fgp_fit_batch: no worker context could be created"); return FGP_ERR_CUDA; }
    if (S.next.load() < n_items) { fgp_set_err(ctx, "fgp_fit_batch: workers stopped early"); return FGP_ERR_CUDA; }
    if (S.infraErr.load()) { fgp_set_err(ctx, "fgp_fit_batch: " + S.infraMsg); return S.infraErr.load(); }
    return 0;
}

extern "C" int fgp_fit_batch(fgp_ctx* ctx, int n, int ds, const double* sIn, int df, const int* k, const double* const* fIn,
                             const double* y, int n_cand, const int* ants, double nugget, int n_starts, int n_presample,
                             const double* u01, const long long* u01_off, int dmax, double* fitness, double* nll,
                             double* sig2, double* theta, int* info) {
    return fit_batch_impl(ctx, n, ds, sIn, df, k, fIn, y, n_cand, ants, nugget, n_starts, n_presample, u01, u01_off, dmax, 0, 0,
                          nullptr, fitness, nll, sig2, theta, info);
}

extern "C" int fgp_fit_batch_holdout(fgp_ctx* ctx, int n, int ds, const double* sIn, int df, const int* k,
                                     const double* const* fIn, const double* y, int n_cand, const int* ants, double nugget,
                                     int n_starts, int n_presample, const double* u01, const long long* u01_off, int dmax,
                                     int n_rep, int n_vl, const int* ind_vl, double* fitness, double* nll, double* sig2,
                                     double* theta, int* info) {
    if (n_rep < 1) return FGP_ERR_ARG;
    return fit_batch_impl(ctx, n, ds, sIn, df, k, fIn, y, n_cand, ants, nugget, n_starts, n_presample, u01, u01_off, dmax, n_rep,
                          n_vl, ind_vl, fitness, nll, sig2, theta, info);
}
