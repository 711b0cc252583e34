// host test driver for fungp_b200/csrc/lbfgsb.hpp: minimises a few bounded analytic functions and prints the result
// as JSON lines (compared with scipy's L-BFGS-B by tests/test_lbfgsb_cpp.py)
#include "../fungp_b200/csrc/lbfgsb.hpp"
#include <cstdio>
#include <cstdlib>
using V = std::vector<double>;
static int prob = 0;
static double fun(const V& x) {
    const int n = (int)x.size();
    double f = 0.0;
    if (prob == 0) {          // convex quadratic, minimiser partly outside the box
        for (int i = 0; i < n; ++i) f += (i + 1) * (x[i] - (0.5 * i - 0.3)) * (x[i] - (0.5 * i - 0.3));
        for (int i = 0; i + 1 < n; ++i) f += 0.3 * x[i] * x[i + 1];
    } else if (prob == 1) {   // Rosenbrock chain
        for (int i = 0; i + 1 < n; ++i) f += 100.0 * (x[i + 1] - x[i] * x[i]) * (x[i + 1] - x[i] * x[i]) + (1 - x[i]) * (1 - x[i]);
    } else {                  // likelihood-like: log-sum + inverse terms, minimum at a bound for some variables
        for (int i = 0; i < n; ++i) f += std::log(x[i] + 0.1) * (i % 2 ? 1.0 : -0.4) + 0.5 / (x[i] + 0.05) + 0.02 * (i + 1) * x[i] * x[i];
        f += 0.1 * std::sin(3.0 * x[0] + x[n - 1]);
    }
    return f;
}
int main(int argc, char** argv) {
    prob = argc > 1 ? atoi(argv[1]) : 0;
    const int n = argc > 2 ? atoi(argv[2]) : 4;
    V x(n), lo(n), up(n);
    for (int i = 0; i < n; ++i) { lo[i] = prob == 2 ? 1e-10 : -1.0; up[i] = prob == 2 ? 2.0 + 0.5 * i : 1.5; x[i] = prob == 2 ? 1.0 : (i % 2 ? 0.2 : -0.5); }
    int nev = 0;
    auto fg = [&](const V& xx, double& f, V& g) {
        // the same central finite differences stats::optim uses (ndeps 1e-3, clipped at the box)
        f = fun(xx); ++nev;
        g.resize(n);
        for (int i = 0; i < n; ++i) {
            V a = xx, b = xx;
            a[i] = std::min(xx[i] + 1e-3, up[i]); b[i] = std::max(xx[i] - 1e-3, lo[i]);
            g[i] = (fun(a) - fun(b)) / (a[i] - b[i]);
        }
        return true;
    };
    auto r = fgp_lbfgsb::minimize(fg, x, lo, up);
    printf("{\"f\": %.17g, \"iters\": %d, \"nfev\": %d, \"conv\": %d, \"x\": [", r.f, r.iterations, r.nfev, r.convergence);
    for (int i = 0; i < n; ++i) printf("%s%.17g", i ? ", " : "", x[i]);
    printf("]}\n");
    return 0;
}
