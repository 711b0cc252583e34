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

// minimise fg over the box [lo, up]; x holds the start and returns the solution
inline Result minimize(const FG& fg, std::vector<double>& x, const std::vector<double>& lo, const std::vector<double>& up,
                       int lmm = 5, double factr = 1e7, double pgtol = 0.0, int maxit = 100) {
    using namespace detail;
    const int n = (int)x.size();
    const double epsmch = std::numeric_limits<double>::epsilon();
    const double tol = factr * epsmch;
    const double INF = std::numeric_limits<double>::infinity();
    Result res;
    for (int i = 0; i < n; ++i) x[i] = std::min(std::max(x[i], lo[i]), up[i]);   // project the start onto the box
    std::vector<double> g(n), xold(n), gold(n), d(n), xc(n), z(n), t(n);
    std::vector<std::vector<double>> S, Y;     // correction pairs, oldest first
    double theta = 1.0, f = 0.0;
    if (!fg(x, f, g)) { res.failed_eval = true; res.convergence = 52; return res; }
    res.nfev = 1;

    auto proj_grad_norm = [&]() {
        double nrm = 0.0;
        for (int i = 0; i < n; ++i) {
            double gi = g[i];
            if (gi < 0.0) gi = std::max(x[i] - up[i], gi);
            else gi = std::min(x[i] - lo[i], gi);
            nrm = std::max(nrm, std::fabs(gi));
        }
        return nrm;
    };
    if (proj_grad_norm() <= pgtol) { res.f = f; return res; }

    // middle matrix M^-1 = [[-D, L'], [L, theta S'S]] (2col x 2col), rows of W = [Y theta*S]
    std::vector<double> Minv;
    int col = 0;
    auto build_middle = [&]() {
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
    };
    auto wrow = [&](int i, std::vector<double>& w) {     // i-th row of W
        w.resize(2 * col);
        for (int j = 0; j < col; ++j) { w[j] = Y[j][i]; w[col + j] = theta * S[j][i]; }
    };
    auto Mtimes = [&](std::vector<double> v) {           // M v  (solve with M^-1)
        if (col) solve_small(Minv, 2 * col, v);
        return v;
    };

    for (int iter = 0;; ++iter) {
        if (iter >= maxit) { res.convergence = 1; break; }
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
        std::vector<double> xbar = xc;
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
        bool restart = false, abnormal = false;
        double stp = 1.0, fold = f, gdold = gd;
        xold = x; gold = g;
        if (gd >= 0.0 || dnorm2 == 0.0) {
            // not a descent direction: drop the memory and restart from steepest descent, or stop
            if (col == 0) abnormal = true; else restart = true;
        } else {
            MTSearch ls;
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
            if (stp <= 0.0 || ls.start(f, gd, stp) == MTSearch::ERROR) {
                if (col == 0) abnormal = true; else restart = true;
            } else {
                int ifun = 0;
                for (;;) {
                    for (int i = 0; i < n; ++i) x[i] = (stp == 1.0) ? xbar[i] : xold[i] + stp * d[i];
                    if (!fg(x, f, g)) { res.failed_eval = true; res.convergence = 52; res.f = fold; x = xold; return res; }
                    ++res.nfev; ++ifun;
                    gd = 0.0;
                    for (int i = 0; i < n; ++i) gd += g[i] * d[i];
                    const MTSearch::Status st = ls.advance(f, gd, stp);
                    if (st == MTSearch::CONVERGED || st == MTSearch::WARNING) break;
                    if (ifun >= 20) {     // too many trial points
                        x = xold; g = gold; f = fold;
                        if (col == 0) abnormal = true; else restart = true;
                        break;
                    }
                }
            }
        }
        if (abnormal) { res.convergence = 52; break; }
        if (restart) { S.clear(); Y.clear(); theta = 1.0; build_middle(); --iter; if (++res.iterations > 2 * maxit) { res.convergence = 52; break; } continue; }
        ++res.iterations;
        // ---------------- convergence tests
        if (proj_grad_norm() <= pgtol) break;
        const double ddum = std::max({std::fabs(fold), std::fabs(f), 1.0});
        if (fold - f <= tol * ddum) break;
        // ---------------- BFGS update
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
    }
    res.f = f;
    return res;
}

}  // namespace fgp_lbfgsb
