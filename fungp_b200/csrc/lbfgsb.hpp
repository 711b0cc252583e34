// L-BFGS-B for the batched candidate fits (fgp_fit_batch): a from-scratch host implementation of the algorithm of
// Byrd, Lu, Nocedal & Zhu (SIAM J. Sci. Comput. 16, 1995) -- generalized Cauchy point, direct primal subspace
// minimisation with backtracking to the box, More'-Thuente line search -- with the controls stats::optim passes
// (R/3_training_SF.R:125: method "L-BFGS-B"; optim defaults lmm = 5, factr = 1e7, pgtol = 0, maxit = 100).
// The reference leaves this to R's own optim(); inside R the fit keeps using optim through `gr=`.
// This copy exists so that MANY candidate structures can be fitted without a round trip to the interpreter per
// function value.  Function and gradient come from one callback (one batched device launch sequence evaluates the
// point and its 2d finite-difference neighbours).
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <vector>

namespace fgp_lbfgsb {

struct Result {
    double f = 0.0;
    int iterations = 0;
    int nfev = 0;
    int convergence = 0;    // 0 converged, 1 iteration limit, 51 warning, 52 abnormal termination (optim's codes)
    bool failed_eval = false;
};

// callback: evaluate f and g at x; return false when the function cannot be evaluated (not positive definite)
using FG = std::function<bool(const std::vector<double>& x, double& f, std::vector<double>& g)>;

namespace detail {

// solve A z = b for a small dense system (k <= 2*lmm) by Gaussian elimination with partial pivoting; A is overwritten
inline bool solve_small(std::vector<double> A, int k, std::vector<double>& b) {
    for (int c = 0; c < k; ++c) {
        int piv = c;
        for (int r = c + 1; r < k; ++r)
            if (std::fabs(A[r * k + c]) > std::fabs(A[piv * k + c])) piv = r;
        if (A[piv * k + c] == 0.0) return false;
        if (piv != c) {
            for (int j = 0; j < k; ++j) std::swap(A[c * k + j], A[piv * k + j]);
            std::swap(b[c], b[piv]);
        }
        for (int r = c + 1; r < k; ++r) {
            const double m = A[r * k + c] / A[c * k + c];
            if (m == 0.0) continue;
            for (int j = c; j < k; ++j) A[r * k + j] -= m * A[c * k + j];
            b[r] -= m * b[c];
        }
    }
    for (int c = k - 1; c >= 0; --c) {
        for (int j = c + 1; j < k; ++j) b[c] -= A[c * k + j] * b[j];
        b[c] /= A[c * k + c];
    }
    return true;
}

// More'-Thuente safeguarded step (the cubic / quadratic interpolation cases of the line search)
inline void mt_step(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp,
                    double dp, bool& brackt, double stpmin, double stpmax) {
    const double sgnd = dp * (dx / std::fabs(dx));
    double stpf, stpc, stpq;
    if (fp > fx) {   // case 1: higher function value -> minimum bracketed
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = std::max({std::fabs(theta), std::fabs(dx), std::fabs(dp)});
        double gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma = -gamma;
        const double p = (gamma - dx) + theta, q = ((gamma - dx) + gamma) + dp, r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        stpf = (std::fabs(stpc - stx) < std::fabs(stpq - stx)) ? stpc : stpc + (stpq - stpc) / 2.0;
        brackt = true;
    } else if (sgnd < 0.0) {   // case 2: lower value, derivatives of opposite sign -> bracketed
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = std::max({std::fabs(theta), std::fabs(dx), std::fabs(dp)});
        double gamma = s * std::sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dx, r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
        brackt = true;
    } else if (std::fabs(dp) < std::fabs(dx)) {   // case 3: lower value, same sign, derivative magnitude decreases
        const double theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        const double s = std::max({std::fabs(theta), std::fabs(dx), std::fabs(dp)});
        double gamma = s * std::sqrt(std::max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        const double p = (gamma - dp) + theta, q = (gamma + (dx - dp)) + gamma, r = p / q;
        if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
        else stpc = (stp > stx) ? stpmax : stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            stpf = (std::fabs(stpc - stp) < std::fabs(stpq - stp)) ? stpc : stpq;
            stpf = (stp > stx) ? std::min(stp + 0.66 * (sty - stp), stpf) : std::max(stp + 0.66 * (sty - stp), stpf);
        } else {
            stpf = (std::fabs(stpc - stp) > std::fabs(stpq - stp)) ? stpc : stpq;
            stpf = std::min(stpmax, stpf);
            stpf = std::max(stpmin, stpf);
        }
    } else {   // case 4: lower value, same sign, derivative does not decrease
        if (brackt) {
            const double theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            const double s = std::max({std::fabs(theta), std::fabs(dy), std::fabs(dp)});
            double gamma = s * std::sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            const double p = (gamma - dp) + theta, q = ((gamma - dp) + gamma) + dy, r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else stpf = (stp > stx) ? stpmax : stpmin;
    }
    if (fp > fx) { sty = stp; fy = fp; dy = dp; }
    else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

// More'-Thuente line search in reverse communication: call start() once, then advance() after every evaluation
struct MTSearch {
    double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0, stpmax = 0.0;
    bool brackt = false;
    int stage = 1;
    double finit = 0, ginit = 0, gtest = 0, width = 0, width1 = 0;
    double stx = 0, fx = 0, gx = 0, sty = 0, fy = 0, gy = 0, stmin = 0, stmax = 0;
    enum Status { FG_NEEDED, CONVERGED, WARNING, ERROR };
    Status start(double f, double g, double& stp) {
        if (stp < stpmin || stp > stpmax || g >= 0.0) return ERROR;
        brackt = false; stage = 1; finit = f; ginit = g; gtest = ftol * ginit;
        width = stpmax - stpmin; width1 = width / 0.5;
        stx = 0; fx = finit; gx = ginit; sty = 0; fy = finit; gy = ginit;
        stmin = 0; stmax = stp + 4.0 * stp;
        return FG_NEEDED;
    }
    Status advance(double f, double g, double& stp) {
        const double ftest = finit + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        Status st = FG_NEEDED;
        if (brackt && (stp <= stmin || stp >= stmax)) st = WARNING;                 // rounding errors
        if (brackt && stmax - stmin <= xtol * stmax) st = WARNING;                  // xtol
        if (stp == stpmax && f <= ftest && g <= gtest) st = WARNING;                // at stpmax
        if (stp == stpmin && (f > ftest || g >= gtest)) st = WARNING;               // at stpmin
        if (f <= ftest && std::fabs(g) <= gtol * (-ginit)) st = CONVERGED;
        if (st != FG_NEEDED) return st;
        if (stage == 1 && f <= fx && f > ftest) {
            double fm = f - stp * gtest, fxm = fx - stx * gtest, fym = fy - sty * gtest;
            double gm = g - gtest, gxm = gx - gtest, gym = gy - gtest;
            mt_step(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
            fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
        } else mt_step(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax);
        if (brackt) {
            if (std::fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
            width1 = width; width = std::fabs(sty - stx);
        }
        if (brackt) { stmin = std::min(stx, sty); stmax = std::max(stx, sty); }
        else { stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx); }
        stp = std::max(stp, stpmin);
        stp = std::min(stp, stpmax);
        if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
        return FG_NEEDED;
    }
};

}  // namespace detail

// L-BFGS-B in reverse communication: the caller evaluates f and the gradient wherever the machine asks.
//   Machine m; m.init(x0, lo, up);  while (!m.done) { evaluate at m.x; m.feed(ok, f, g); }   -> m.x, m.res
// This is what lets ONE host thread advance the fits of many candidate structures in lock-step (fitlock.cu): every
// round the pending points of all machines go to the device in a single batched launch sequence.  minimize() below is
// the same machine driven by a callback, so both callers run identical arithmetic.
struct Machine {
    Result res;
    std::vector<double> x;          // where f and g are needed next; the solution once `done`
    bool done = false;

    void init(const std::vector<double>& x0, const std::vector<double>& lo_, const std::vector<double>& up_, int lmm_ = 5,
              double factr = 1e7, double pgtol_ = 0.0, int maxit_ = 100) {
        n = (int)x0.size();
        lo = lo_; up = up_; lmm = lmm_; pgtol = pgtol_; maxit = maxit_;
        epsmch = std::numeric_limits<double>::epsilon();
        tol = factr * epsmch;
        res = Result();
        done = false;
        x = x0;
        for (int i = 0; i < n; ++i) x[i] = std::min(std::max(x[i], lo[i]), up[i]);   // project the start onto the box
        g.assign(n, 0.0); xold.assign(n, 0.0); gold.assign(n, 0.0); d.assign(n, 0.0); xc.assign(n, 0.0); t.assign(n, 0.0);
        S.clear(); Y.clear(); Minv.clear();
        theta = 1.0; f = 0.0; col = 0; iter = 0;
        state = ST_INIT;
    }

    // the evaluation at x: ok = false when the function cannot be evaluated there (not positive definite)
    void feed(bool ok, double fnew, const std::vector<double>& gnew) {
        using namespace detail;
        if (done) return;
        if (state == ST_INIT) {
            if (!ok) { res.failed_eval = true; res.convergence = 52; done = true; return; }
            f = fnew; g = gnew;
            res.nfev = 1;
            if (proj_grad_norm() <= pgtol) { res.f = f; done = true; return; }
            next_direction();
            return;
        }
        // ---- line search trial point
        if (!ok) { res.failed_eval = true; res.convergence = 52; res.f = fold; x = xold; done = true; return; }
        f = fnew; g = gnew;
        ++res.nfev; ++ifun;
        double gd = 0.0;
        for (int i = 0; i < n; ++i) gd += g[i] * d[i];
        const MTSearch::Status st = ls.advance(f, gd, stp);
        if (st == MTSearch::CONVERGED || st == MTSearch::WARNING) { after_line_search(); return; }
        if (ifun >= 20) {     // too many trial points
            x = xold; g = gold; f = fold;
            if (col == 0) { res.convergence = 52; finish(); return; }
            restart_memory();
            return;
        }
        trial_point();
    }

private:
    enum { ST_INIT, ST_LS } state = ST_INIT;
    int n = 0, lmm = 5, maxit = 100, col = 0, iter = 0, ifun = 0;
    double epsmch = 0, tol = 0, pgtol = 0, theta = 1.0, f = 0, fold = 0, gdold = 0, stp = 1.0;
    std::vector<double> lo, up, g, xold, gold, d, xc, t, xbar, Minv;
    std::vector<std::vector<double>> S, Y;     // correction pairs, oldest first
    detail::MTSearch ls;

    void finish() { res.f = f; done = true; }
    double proj_grad_norm() const {
        double nrm = 0.0;
        for (int i = 0; i < n; ++i) {
            double gi = g[i];
            if (gi < 0.0) gi = std::max(x[i] - up[i], gi);
            else gi = std::min(x[i] - lo[i], gi);
            nrm = std::max(nrm, std::fabs(gi));
        }
        return nrm;
    }
    // middle matrix M^-1 = [[-D, L'], [L, theta S'S]] (2col x 2col), rows of W = [Y theta*S]
    void build_middle() {
        col = (int)S.size();
        const int k = 2 * col;
        Minv.assign((size_t)k * k, 0.0);
        for (int i = 0; i < col; ++i)
            for (int j = 0; j < col; ++j) {
                double sy = 0.0, ss = 0.0;
                for (int r = 0; r < n; ++r) { sy += S[i][r] * Y[j][r]; ss += S[i][r] * S[j][r]; }
                if (i == j) Minv[i * k + j] = -sy;                        // -D
                if (i > j) { Minv[(col + i) * k + j] = sy; Minv[j * k + (col + i)] = sy; }   // L and L'
                Minv[(col + i) * k + (col + j)] = theta * ss;
            }
    }
    void wrow(int i, std::vector<double>& w) const {     // i-th row of W
        w.resize(2 * col);
        for (int j = 0; j < col; ++j) { w[j] = Y[j][i]; w[col + j] = theta * S[j][i]; }
    }
    std::vector<double> Mtimes(std::vector<double> v) const {           // M v  (solve with M^-1)
        if (col) detail::solve_small(Minv, 2 * col, v);
        return v;
    }
    void restart_memory() {
        // not a descent direction / failed search: drop the memory and restart from steepest descent
        S.clear(); Y.clear(); theta = 1.0; build_middle();
        if (++res.iterations > 2 * maxit) { res.convergence = 52; finish(); return; }
        next_direction();
    }
    void trial_point() {
        for (int i = 0; i < n; ++i) x[i] = (stp == 1.0) ? xbar[i] : xold[i] + stp * d[i];
        state = ST_LS;
    }

    // generalized Cauchy point, subspace minimisation, start of the line search: leaves the first trial point in x
    void next_direction() {
        using namespace detail;
        const double INF = std::numeric_limits<double>::infinity();
        if (iter >= maxit) { res.convergence = 1; finish(); return; }
        // ---------------- generalized Cauchy point
        const int k = 2 * col;
        std::vector<double> p(k, 0.0), c(k, 0.0), wb;
        std::vector<int> order;
        for (int i = 0; i < n; ++i) {
            if (g[i] < 0.0) t[i] = (x[i] - up[i]) / g[i];
            else if (g[i] > 0.0) t[i] = (x[i] - lo[i]) / g[i];
            else t[i] = INF;
            d[i] = (t[i] <= 0.0) ? 0.0 : -g[i];
            xc[i] = x[i];
            if (t[i] > 0.0 && t[i] < INF) order.push_back(i);
        }
        std::sort(order.begin(), order.end(), [&](int a, int b) { return t[a] < t[b]; });
        double fp = 0.0;
        for (int i = 0; i < n; ++i) {
            fp -= d[i] * d[i];
            if (d[i] != 0.0 && col) { wrow(i, wb); for (int j = 0; j < k; ++j) p[j] += wb[j] * d[i]; }
        }
        double fpp = -theta * fp;
        if (col) { auto Mp = Mtimes(p); for (int j = 0; j < k; ++j) fpp -= p[j] * Mp[j]; }
        const double fpp_org = fpp;
        double dtm = (fpp != 0.0) ? -fp / fpp : 0.0, told = 0.0;
        bool all_fixed = (fp == 0.0);
        size_t ib = 0;
        if (!all_fixed) {
            while (ib < order.size()) {
                const int b = order[ib];
                const double dt = t[b] - told;
                if (dtm < dt) break;
                // variable b reaches its bound
                xc[b] = (d[b] > 0.0) ? up[b] : lo[b];
                const double zb = xc[b] - x[b], gb = g[b];
                for (int j = 0; j < k; ++j) c[j] += dt * p[j];
                fp += dt * fpp + gb * gb + theta * gb * zb;
                fpp -= theta * gb * gb;
                if (col) {
                    wrow(b, wb);
                    auto Mc = Mtimes(c), Mp = Mtimes(p), Mw = Mtimes(wb);
                    double wMc = 0, wMp = 0, wMw = 0;
                    for (int j = 0; j < k; ++j) { wMc += wb[j] * Mc[j]; wMp += wb[j] * Mp[j]; wMw += wb[j] * Mw[j]; }
                    fp -= gb * wMc;
                    fpp -= 2.0 * gb * wMp + gb * gb * wMw;
                    for (int j = 0; j < k; ++j) p[j] += gb * wb[j];
                }
                d[b] = 0.0;
                told = t[b];
                ++ib;
                fpp = std::max(epsmch * fpp_org, fpp);     // the quadratic model stays convex along the path
                dtm = -fp / fpp;
            }
            if (dtm < 0.0) dtm = 0.0;
            told += dtm;
            for (int i = 0; i < n; ++i)
                if (d[i] != 0.0) xc[i] = x[i] + told * d[i];
            for (int j = 0; j < k; ++j) c[j] += dtm * p[j];
        }
        // ---------------- subspace minimisation over the variables that are free at the Cauchy point
        xbar = xc;
        std::vector<int> freev;
        for (int i = 0; i < n; ++i)
            if (xc[i] > lo[i] && xc[i] < up[i]) freev.push_back(i);
        if (col > 0 && !freev.empty()) {
            const int nf = (int)freev.size();
            auto Mc = Mtimes(c);
            std::vector<double> r(nf);
            for (int a = 0; a < nf; ++a) {
                const int i = freev[a];
                wrow(i, wb);
                double wMc = 0.0;
                for (int j = 0; j < k; ++j) wMc += wb[j] * Mc[j];
                r[a] = g[i] + theta * (xc[i] - x[i]) - wMc;
            }
            // v = M W'Z r ;  N = I - (1/theta) M W'ZZ'W ;  v = N^-1 v ;  du = -(1/theta) r - (1/theta^2) Z'W v
            std::vector<double> v(k, 0.0), WtZZtW((size_t)k * k, 0.0);
            for (int a = 0; a < nf; ++a) {
                wrow(freev[a], wb);
                for (int j = 0; j < k; ++j) {
                    v[j] += wb[j] * r[a];
                    for (int l = 0; l < k; ++l) WtZZtW[(size_t)j * k + l] += wb[j] * wb[l];
                }
            }
            v = Mtimes(v);
            std::vector<double> N((size_t)k * k, 0.0);
            for (int l = 0; l < k; ++l) {      // column l of M * WtZZtW
                std::vector<double> colv(k);
                for (int j = 0; j < k; ++j) colv[j] = WtZZtW[(size_t)j * k + l];
                colv = Mtimes(colv);
                for (int j = 0; j < k; ++j) N[(size_t)j * k + l] = (j == l ? 1.0 : 0.0) - colv[j] / theta;
            }
            bool ok = solve_small(N, k, v);
            if (ok) {
                std::vector<double> du(nf);
                double alpha = 1.0;
                for (int a = 0; a < nf; ++a) {
                    const int i = freev[a];
                    wrow(i, wb);
                    double wv = 0.0;
                    for (int j = 0; j < k; ++j) wv += wb[j] * v[j];
                    du[a] = -r[a] / theta - wv / (theta * theta);
                    if (du[a] > 0.0) alpha = std::min(alpha, (up[i] - xc[i]) / du[a]);
                    else if (du[a] < 0.0) alpha = std::min(alpha, (lo[i] - xc[i]) / du[a]);
                }
                alpha = std::max(alpha, 0.0);
                for (int a = 0; a < nf; ++a) {
                    const int i = freev[a];
                    xbar[i] = std::min(std::max(xc[i] + alpha * du[a], lo[i]), up[i]);
                }
            }
        }
        // ---------------- line search along d = xbar - x
        double dnorm2 = 0.0, gd = 0.0;
        for (int i = 0; i < n; ++i) { d[i] = xbar[i] - x[i]; dnorm2 += d[i] * d[i]; gd += g[i] * d[i]; }
        stp = 1.0; fold = f; gdold = gd;
        xold = x; gold = g;
        bool bad = false;
        if (gd >= 0.0 || dnorm2 == 0.0) bad = true;
        else {
            ls = MTSearch();
            double stpmx = 1.0;
            if (iter != 0) {
                stpmx = 1e10;
                for (int i = 0; i < n; ++i) {
                    if (d[i] < 0.0) { const double a2 = lo[i] - x[i]; if (a2 >= 0.0) stpmx = 0.0; else if (d[i] * stpmx < a2) stpmx = a2 / d[i]; }
                    else if (d[i] > 0.0) { const double a2 = up[i] - x[i]; if (a2 <= 0.0) stpmx = 0.0; else if (d[i] * stpmx > a2) stpmx = a2 / d[i]; }
                }
            }
            ls.stpmax = stpmx;
            stp = std::min(1.0, stpmx);
            if (stp <= 0.0 || ls.start(f, gd, stp) == MTSearch::ERROR) bad = true;
        }
        if (bad) {
            if (col == 0) { res.convergence = 52; finish(); return; }      // abnormal termination
            restart_memory();
            return;
        }
        ifun = 0;
        trial_point();
    }

    // the line search has accepted x: convergence tests, BFGS update, next iteration
    void after_line_search() {
        ++res.iterations;
        if (proj_grad_norm() <= pgtol) { finish(); return; }
        const double ddum = std::max({std::fabs(fold), std::fabs(f), 1.0});
        if (fold - f <= tol * ddum) { finish(); return; }
        std::vector<double> s(n), y(n);
        double sy = 0.0, yy = 0.0;
        for (int i = 0; i < n; ++i) { s[i] = x[i] - xold[i]; y[i] = g[i] - gold[i]; sy += s[i] * y[i]; yy += y[i] * y[i]; }
        const double ddum2 = -gdold * stp;
        if (sy > epsmch * ddum2) {
            if ((int)S.size() == lmm) { S.erase(S.begin()); Y.erase(Y.begin()); }
            S.push_back(s); Y.push_back(y);
            theta = yy / sy;
            build_middle();
        }
        ++iter;
        next_direction();
    }
};

// minimise fg over the box [lo, up]; x holds the start and returns the solution
inline Result minimize(const FG& fg, std::vector<double>& x, const std::vector<double>& lo, const std::vector<double>& up,
                       int lmm = 5, double factr = 1e7, double pgtol = 0.0, int maxit = 100) {
    Machine m;
    m.init(x, lo, up, lmm, factr, pgtol, maxit);
    std::vector<double> g(x.size());
    while (!m.done) {
        double f = 0.0;
        const bool ok = fg(m.x, f, g);
        m.feed(ok, f, g);
    }
    x = m.x;
    return m.res;
}

}  // namespace fgp_lbfgsb
