// fgp_fit_batch: batched scoring of candidate model structures for the ant-colony search.
// One candidate = what fitNtests_ACO does per ant (R/3_ant_admin.R:810-831): formatSol_ACO decode (:301-373) ->
// fgpm(): dimReduction, distances, setBounds / setSPoints / optimHypers (R/3_training_SF.R:57-223), preMats
// (R/4_prediction_SF.R:24-32) -> getFitness: Q2 of the leave-one-out predictions (R/1_Xfgpm_Class.R:522-543).
// Host orchestration only: every number comes from the same C-ABI entry points a single fgpm() uses
// (fgp_project, fgp_problem_create, fgp_dist_max, fgp_nll_batch, fgp_premats, fgp_loocv); the optimiser is the host
// L-BFGS-B of lbfgsb.hpp with one batched device call per iterate (value + 2d finite-difference neighbours).
// Candidates are independent: they are dealt from a queue to worker threads, `fit_workers` per GPU of the context,
// each with its own streams and workspaces, so small-n fits overlap on the chip and GPUs never exchange data.
#include "ctx.h"
#include "lbfgsb.hpp"
#include "../../include/fungp_b200.h"
#include <atomic>
#include <cstring>
#include <map>
#include <thread>
#include <tuple>

namespace {

struct Proj { std::vector<double> B, J, coefs; int pe = 0; bool ok = false; };

struct Shared {
    int n, ds, df;
    const double* sIn; const int* k; const double* const* fIn; const double* y;
    int n_cand, antLen; const int* ants;
    double nugget; int n_starts, n_presample;
    const double* u01; const long long* u01_off; int dmax;
    double *fitness, *nll, *sig2, *theta; int* info;
    std::mutex mu;
    std::map<std::tuple<int, int, int>, Proj> cache;    // (input, p, basis) -> projection, shared by all candidates
    std::atomic<int> next{0};
    std::atomic<int> failures{0};
};

const Proj& projection(Shared& S, int j, int p, int basis) {
    std::lock_guard<std::mutex> lk(S.mu);
    auto key = std::make_tuple(j, p, basis);
    auto it = S.cache.find(key);
    if (it != S.cache.end()) return it->second;
    Proj pr;
    const int kk = S.k[j];
    pr.pe = p ? p : kk;
    pr.B.resize((size_t)kk * pr.pe); pr.J.resize((size_t)pr.pe * pr.pe); pr.coefs.resize((size_t)S.n * pr.pe);
    pr.ok = fgp_project(nullptr, S.fIn[j], S.n, kk, p, basis, nullptr, pr.B.data(), pr.J.data(), pr.coefs.data()) == 0;
    return S.cache.emplace(key, std::move(pr)).first->second;
}

void crash(Shared& S, int c, int code) {
    S.fitness[c] = std::nan(""); S.nll[c] = std::nan(""); S.sig2[c] = std::nan("");
    for (int i = 0; i < S.dmax; ++i) S.theta[(size_t)c * S.dmax + i] = std::nan("");
    S.info[c] = code;
}

void fit_one(Shared& S, fgp_ctx* ctx, int c) {
    const int* ant = S.ants + (size_t)c * S.antLen;
    // ---- formatSol_ACO
    std::vector<double> sAct;
    int dsA = 0;
    for (int i = 0; i < S.ds; ++i)
        if (ant[i] == 1) {
            sAct.insert(sAct.end(), S.sIn + (size_t)i * S.n, S.sIn + (size_t)(i + 1) * S.n);
            ++dsA;
        }
    std::vector<const double*> coefs, Js;
    std::vector<int> p, dis;
    for (int j = 0; j < S.df; ++j) {
        const int* a = ant + S.ds + 4 * j;
        if (a[0] != 1) continue;
        const int dist = a[1], dim = a[2], bas = (a[3] == FGP_BASIS_PCA) ? FGP_BASIS_PCA : FGP_BASIS_BSPLINES;
        if ((dist != FGP_DIST_BYGROUP && dist != FGP_DIST_BYINDEX) || dim < 0 || dim > S.k[j]) { crash(S, c, -1); return; }
        const Proj& pr = projection(S, j, dim, bas);
        if (!pr.ok) { crash(S, c, -2); return; }
        coefs.push_back(pr.coefs.data()); Js.push_back(pr.J.data()); p.push_back(pr.pe); dis.push_back(dist);
    }
    const int ker = ant[S.antLen - 1];
    const int dfA = (int)coefs.size();
    if ((dsA == 0 && dfA == 0) || ker < 1 || ker > 3) { crash(S, c, -1); return; }

    fgp_problem* P = nullptr;
    if (fgp_problem_create(ctx, S.n, dsA, dsA ? sAct.data() : nullptr, dfA, p.data(), dis.data(), coefs.data(), Js.data(), S.y, &P)) {
        crash(S, c, -3); return;
    }
    const int d = fgp_problem_nls(P);
    auto done = [&](int code) { if (code) crash(S, c, code); fgp_problem_destroy(P); };
    if (d > S.dmax) { done(-4); return; }
    // ---- setBounds_*: lower 1e-10, upper 2 max(D_l)
    std::vector<double> lo(d, 1e-10), up(d);
    if (fgp_dist_max(P, up.data())) { done(-3); return; }
    for (int i = 0; i < d; ++i) up[i] *= 2.0;
    // ---- setSPoints_*: the presample draws were made by the caller (R's runif, in ant order)
    const int npre = std::max(S.n_presample, S.n_starts);
    const double* u = S.u01 + S.u01_off[c];
    std::vector<double> pts((size_t)d * npre), v(npre), s2(npre);
    std::vector<int> inf(npre);
    for (int j = 0; j < npre; ++j)
        for (int i = 0; i < d; ++i) pts[i + (size_t)j * d] = lo[i] + u[i + (size_t)j * d] * (up[i] - lo[i]);
    if (fgp_nll_batch(P, ker, S.nugget, npre, pts.data(), nullptr, v.data(), s2.data(), inf.data())) { done(-3); return; }
    for (int j = 0; j < npre; ++j)
        if (inf[j]) { done(inf[j]); return; }       // quirk Q6: a chol failure while presampling aborts the fit
    std::vector<int> ord(npre);
    for (int j = 0; j < npre; ++j) ord[j] = j;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return v[a] < v[b]; });
    // ---- optimHypers_*: L-BFGS-B from each start, keep the best
    const int nfd = 2 * d + 1;
    std::vector<double> fdp((size_t)d * nfd), fv(nfd), fs(nfd);
    std::vector<int> fi(nfd);
    fgp_lbfgsb::FG fg = [&](const std::vector<double>& x, double& f, std::vector<double>& g) {
        for (int j = 0; j < nfd; ++j)
            for (int i = 0; i < d; ++i) fdp[i + (size_t)j * d] = x[i];
        for (int i = 0; i < d; ++i) {
            fdp[i + (size_t)(1 + 2 * i) * d] = std::min(x[i] + 1e-3, up[i]);
            fdp[i + (size_t)(2 + 2 * i) * d] = std::max(x[i] - 1e-3, lo[i]);
        }
        if (fgp_nll_batch(P, ker, S.nugget, nfd, fdp.data(), nullptr, fv.data(), fs.data(), fi.data())) return false;
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
        if (r.failed_eval && r.nfev <= 1) continue;          // tryCatch per start (R/3_training_SF.R:141-153)
        if (!have || r.f < bestf) { have = true; bestf = r.f; bestx = x; }
    }
    if (!have) { done(-5); return; }                        // "All model optimizations crashed!"
    // ---- variance at the optimum, preMats, Q2loocv
    double f1, s1; int i1;
    if (fgp_nll_batch(P, ker, S.nugget, 1, bestx.data(), nullptr, &f1, &s1, &i1) || i1) { done(i1 ? i1 : -3); return; }
    int rc = fgp_premats(P, ker, S.nugget, s1, bestx.data(), nullptr, nullptr);
    if (rc) { done(rc); return; }
    double q2 = 0.0;
    rc = fgp_loocv(P, nullptr, &q2);
    if (rc) { done(rc); return; }
    S.fitness[c] = q2; S.nll[c] = bestf; S.sig2[c] = s1; S.info[c] = 0;
    for (int i = 0; i < S.dmax; ++i) S.theta[(size_t)c * S.dmax + i] = i < d ? bestx[i] : std::nan("");
    done(0);
}

}  // namespace

extern "C" int fgp_fit_batch(fgp_ctx* ctx, int n, int ds, const double* sIn, int df, const int* k, const double* const* fIn,
                             const double* y, int n_cand, const int* ants, double nugget, int n_starts, int n_presample,
                             const double* u01, const long long* u01_off, int dmax, double* fitness, double* nll,
                             double* sig2, double* theta, int* info) {
    if (!ctx || n <= 0 || ds < 0 || df < 0 || (ds > 0 && !sIn) || (df > 0 && (!k || !fIn)) || !y || n_cand < 0 || !ants ||
        n_starts < 1 || n_presample < 1 || !u01 || !u01_off || dmax < 1 || !fitness || !nll || !sig2 || !theta || !info)
        return FGP_ERR_ARG;
    Shared S;
    S.n = n; S.ds = ds; S.df = df; S.sIn = sIn; S.k = k; S.fIn = fIn; S.y = y;
    S.n_cand = n_cand; S.antLen = ds + 4 * df + 1; S.ants = ants;
    S.nugget = nugget; S.n_starts = n_starts; S.n_presample = n_presample;
    S.u01 = u01; S.u01_off = u01_off; S.dmax = dmax;
    S.fitness = fitness; S.nll = nll; S.sig2 = sig2; S.theta = theta; S.info = info;
    const int nd = (int)ctx->dev.size();
    const int wpd = std::max(1, ctx->fit_workers);
    std::vector<std::thread> th;
    std::atomic<int> ctxFail{0};
    for (int dI = 0; dI < nd; ++dI)
        for (int w = 0; w < wpd; ++w) {
            const int device = ctx->dev[dI].device;
            th.emplace_back([&, device]() {
                fgp_ctx* wc = nullptr;
                if (fgp_ctx_create(1, &device, &wc) != 0) { ctxFail++; return; }
                wc->streams = ctx->streams; wc->outer = ctx->outer; wc->lookahead = ctx->lookahead;
                for (;;) {
                    const int c = S.next.fetch_add(1);
                    if (c >= S.n_cand) break;
                    fit_one(S, wc, c);
                }
                fgp_ctx_destroy(wc);
            });
        }
    for (auto& t : th) t.join();
    cudaSetDevice(ctx->dev[0].device);
    if (ctxFail.load() == nd * wpd && n_cand > 0) { fgp_set_err(ctx, "fgp_fit_batch: no worker context could be created"); return FGP_ERR_CUDA; }
    if (S.next.load() < S.n_cand) { fgp_set_err(ctx, "fgp_fit_batch: workers stopped early"); return FGP_ERR_CUDA; }
    return 0;
}
