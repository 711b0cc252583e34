// Shared by the two candidate-scoring drivers (fitbatch.cu: one host thread per candidate; fitlock.cu: one host thread
// per GPU advancing many candidates in lock-step): the call's arguments, the projection cache and the decoding of a
// candidate structure (formatSol_ACO, R/3_ant_admin.R:301-373) into the arguments of fgp_problem_create.
#pragma once
#include "ctx.h"
#include "../../include/fungp_b200.h"
#include <atomic>
#include <cmath>
#include <map>
#include <tuple>

void whitening_factor(const double* J, int p, std::vector<double>& W);
// z = c W (m x p, column-major): the whitened coefficients of an L2_bygroup input (api.cu, shared with build_coords)
void whiten_columns(const double* c, int m, int p, const double* W, double* z);

namespace fgp_fit {

struct Proj {
    std::vector<double> B, J, coefs, coefsVl;
    int pe = 0;
    bool ok = false;
    // lock-step driver: what a candidate needs from this input without touching the device again
    std::vector<double> W, z;            // whitening factor of J and whitened training coefficients (L2_bygroup)
    std::vector<double> colRange;        // max - min of every raw coefficient column (upper bounds of L2_byindex terms)
    double groupMax = -1.0;              // max L2_bygroup distance between training curves (< 0: not computed yet)
};

struct Shared {
    int n, ds, df;
    const double* sIn; const int* k; const double* const* fIn; const double* y;
    int n_cand, antLen; const int* ants;
    double nugget; int n_starts, n_presample;
    const double* u01; const long long* u01_off; int dmax;
    double *fitness, *nll, *sig2, *theta; int* info;
    // hold-out validation (eval_houtv_ACO): n_rep splits, n_vl validation rows each (1-based row indices, column-major
    // n_vl x n_rep); n_rep = 0 -> leave-one-out fitness on the full data
    int n_rep = 0, n_vl = 0; const int* ind_vl = nullptr;
    std::vector<std::vector<int>> trRows, vlRows;       // per replicate: 0-based training / validation rows
    std::vector<std::vector<double>> yTr, yVl;
    std::mutex mu;
    std::map<std::tuple<int, int, int, int>, Proj> cache;    // (input, p, basis, replicate) -> projection, shared by all items
    std::atomic<int> next{0};
    // infrastructure failures (CUDA error, out of memory: rc < 0 from the library itself) are NOT model crashes: the
    // first one is kept and returned by fgp_fit_batch instead of a generation of NA ants
    std::atomic<int> infraErr{0};
    std::string infraMsg;
    void infra(int rc, fgp_ctx* wc) {
        int expect = 0;
        if (infraErr.compare_exchange_strong(expect, rc)) {
            std::lock_guard<std::mutex> lk(mu);
            infraMsg = wc ? wc->err : std::string();
        }
    }
};

// gather rows of a column-major n x k matrix
inline std::vector<double> take_rows(const double* M, int n, int k, const std::vector<int>& rows) {
    std::vector<double> out((size_t)rows.size() * k);
    for (int c = 0; c < k; ++c)
        for (size_t r = 0; r < rows.size(); ++r) out[r + (size_t)c * rows.size()] = M[rows[r] + (size_t)c * n];
    return out;
}

// dimReduction of input j on the training rows of replicate `rep` (-1: all rows) and, for a hold-out split, the
// projection of the validation curves with the same basis (R/1_fgpm_Class.R:844-845).  The contractions over the curves
// run on the caller's device context; the cache lock is only held to look up / insert.
inline const Proj* projection(Shared& S, fgp_ctx* wc, int j, int p, int basis, int rep) {
    const auto key = std::make_tuple(j, p, basis, rep);
    {
        std::lock_guard<std::mutex> lk(S.mu);
        auto it = S.cache.find(key);
        if (it != S.cache.end()) return &it->second;
    }
    Proj pr;
    const int kk = S.k[j];
    pr.pe = p ? p : kk;
    pr.B.resize((size_t)kk * pr.pe); pr.J.resize((size_t)pr.pe * pr.pe);
    int ntr = S.n;
    if (rep < 0) {
        pr.coefs.resize((size_t)S.n * pr.pe);
        pr.ok = fgp_project(wc, S.fIn[j], S.n, kk, p, basis, nullptr, pr.B.data(), pr.J.data(), pr.coefs.data()) == 0;
    } else {
        const std::vector<double> ftr = take_rows(S.fIn[j], S.n, kk, S.trRows[rep]), fvl = take_rows(S.fIn[j], S.n, kk, S.vlRows[rep]);
        ntr = (int)S.trRows[rep].size();
        const int nvl = (int)S.vlRows[rep].size();
        pr.coefs.resize((size_t)ntr * pr.pe); pr.coefsVl.resize((size_t)nvl * pr.pe);
        std::vector<double> Jtmp((size_t)pr.pe * pr.pe), Btmp((size_t)kk * pr.pe);
        pr.ok = fgp_project(wc, ftr.data(), ntr, kk, p, basis, nullptr, pr.B.data(), pr.J.data(), pr.coefs.data()) == 0 &&
                fgp_project(wc, fvl.data(), nvl, kk, p, basis, p ? pr.B.data() : nullptr, Btmp.data(), Jtmp.data(), pr.coefsVl.data()) == 0;
    }
    if (pr.ok) {
        pr.colRange.resize(pr.pe);
        for (int a = 0; a < pr.pe; ++a) {
            const double* x = &pr.coefs[(size_t)a * ntr];
            double lo = x[0], hi = x[0];
            for (int i = 1; i < ntr; ++i) { lo = std::min(lo, x[i]); hi = std::max(hi, x[i]); }
            pr.colRange[a] = hi - lo;
        }
    }
    std::lock_guard<std::mutex> lk(S.mu);
    return &S.cache.emplace(key, std::move(pr)).first->second;      // a racing insert of the same key keeps the first
}

inline void crash(Shared& S, int c, int code) {
    S.fitness[c] = std::nan(""); S.nll[c] = std::nan(""); S.sig2[c] = std::nan("");
    for (int i = 0; i < S.dmax; ++i) S.theta[(size_t)c * S.dmax + i] = std::nan("");
    S.info[c] = code;
}

// one work item decoded: the arguments fgpm() would receive for this ant (and replicate)
struct Decoded {
    int cand = 0, rep = -1, n = 0, ker = 0, dsA = 0, dfA = 0, d = 0;
    std::vector<int> sIdx;                      // active scalar inputs (columns of sIn)
    std::vector<int> fIdx;                      // active functional inputs
    std::vector<const Proj*> proj;              // their projections
    std::vector<int> p, dis;
    int code = 0;                               // != 0: bad structure (-1) / projection failed (-2)
};

inline Decoded decode_item(Shared& S, fgp_ctx* wc, int item) {
    Decoded D;
    const int nrep = std::max(1, S.n_rep);
    D.cand = item / nrep; D.rep = S.n_rep > 0 ? item % nrep : -1;
    const int* ant = S.ants + (size_t)D.cand * S.antLen;
    D.n = D.rep < 0 ? S.n : (int)S.trRows[D.rep].size();
    for (int i = 0; i < S.ds; ++i)
        if (ant[i] == 1) D.sIdx.push_back(i);
    D.dsA = (int)D.sIdx.size();
    D.d = D.dsA;
    for (int j = 0; j < S.df; ++j) {
        const int* a = ant + S.ds + 4 * j;
        if (a[0] != 1) continue;
        const int dist = a[1], dim = a[2], bas = (a[3] == FGP_BASIS_PCA) ? FGP_BASIS_PCA : FGP_BASIS_BSPLINES;
        if ((dist != FGP_DIST_BYGROUP && dist != FGP_DIST_BYINDEX) || dim < 0 || dim > S.k[j]) { D.code = -1; return D; }
        const Proj* pr = projection(S, wc, j, dim, bas, D.rep);
        if (!pr->ok) { D.code = -2; return D; }
        D.fIdx.push_back(j); D.proj.push_back(pr); D.p.push_back(pr->pe); D.dis.push_back(dist);
        D.d += dist == FGP_DIST_BYINDEX ? pr->pe : 1;
    }
    D.dfA = (int)D.fIdx.size();
    D.ker = ant[S.antLen - 1];
    if ((D.dsA == 0 && D.dfA == 0) || D.ker < 1 || D.ker > 3) D.code = -1;
    return D;
}

// Q2 of predictions against observations (R/1_Xfgpm_Class.R:541, 548)
inline double q2_of(const double* y, const double* yhat, int n) {
    double mu = 0.0, se = 0.0, st = 0.0;
    for (int i = 0; i < n; ++i) mu += y[i];
    mu /= n;
    for (int i = 0; i < n; ++i) { se += (y[i] - yhat[i]) * (y[i] - yhat[i]); st += (y[i] - mu) * (y[i] - mu); }
    return 1.0 - (se / n) / (st / n);
}

}  // namespace fgp_fit

// fitlock.cu: all items of the call on the GPUs of ctx, `order` = dispatch order (longest first)
int fit_lockstep(fgp_ctx* ctx, fgp_fit::Shared& S, const std::vector<int>& order);
