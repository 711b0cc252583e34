// Dimension reduction of functional inputs (host side of the C ABI): B-spline design matrix, PCA basis and the
// projection coefs = t(solve(B'B, B' f')).  Follows R/7_dimRedFunctions.R:7-60; the arithmetic the reference
// delegates to splines::splineDesign, stats::cov, base::eigen and base::solve is restated here (Cox-de Boor
// recursion, cyclic Jacobi, LU with partial pivoting).  These are O(n k p) / O(k^3) with k <= a few hundred --
// negligible next to the O(n^2) / O(n^3) device work, and kept on the host so that the basis is bit-stable.
#include "../../include/fungp_b200.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

// knots of proj_bsplines (R/7_dimRedFunctions.R:41-49)
void bspline_knots(int k, int p, std::vector<double>& t, int& ord) {
    ord = p <= 3 ? p : 4;
    const int n_inner = p - ord + 2, n_outer = ord - 1;
    t.clear();
    for (int i = 0; i < n_outer; ++i) t.push_back(1.0);
    const double step = n_inner > 1 ? ((double)k - 1.0) / (double)(n_inner - 1) : 0.0;
    for (int i = 0; i < n_inner; ++i) t.push_back(i == n_inner - 1 && n_inner > 1 ? (double)k : 1.0 + i * step);
    for (int i = 0; i < n_outer; ++i) t.push_back((double)k);
}

// all non-zero B-splines of order `ord` at x (The NURBS Book A2.2); span = index j with t[j] <= x < t[j+1]
void basis_funs(const std::vector<double>& t, int ord, int span, double x, double* N) {
    const int deg = ord - 1;
    std::vector<double> left(ord), right(ord);
    N[0] = 1.0;
    for (int j = 1; j <= deg; ++j) {
        left[j] = x - t[span + 1 - j];
        right[j] = t[span + j] - x;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            const double den = right[r + 1] + left[j - r];
            const double tmp = den != 0.0 ? N[r] / den : 0.0;
            N[r] = saved + right[r + 1] * tmp;
            saved = left[j - r] * tmp;
        }
        N[j] = saved;
    }
}

// in-place LU with partial pivoting of a p x p column-major matrix
bool lu_factor(std::vector<double>& A, int p, std::vector<int>& piv) {
    piv.resize(p);
    for (int c = 0; c < p; ++c) {
        int best = c;
        for (int r = c + 1; r < p; ++r)
            if (std::fabs(A[r + (size_t)c * p]) > std::fabs(A[best + (size_t)c * p])) best = r;
        piv[c] = best;
        if (A[best + (size_t)c * p] == 0.0) return false;
        if (best != c)
            for (int j = 0; j < p; ++j) std::swap(A[c + (size_t)j * p], A[best + (size_t)j * p]);
        const double d = 1.0 / A[c + (size_t)c * p];
        for (int r = c + 1; r < p; ++r) A[r + (size_t)c * p] *= d;
        for (int j = c + 1; j < p; ++j) {
            const double a = A[c + (size_t)j * p];
            for (int r = c + 1; r < p; ++r) A[r + (size_t)j * p] -= A[r + (size_t)c * p] * a;
        }
    }
    return true;
}
void lu_solve(const std::vector<double>& A, int p, const std::vector<int>& piv, double* b) {
    for (int c = 0; c < p; ++c) std::swap(b[c], b[piv[c]]);   // all row interchanges first (LAPACK getrs order)
    for (int c = 0; c < p; ++c)
        for (int r = c + 1; r < p; ++r) b[r] -= A[r + (size_t)c * p] * b[c];
    for (int c = p - 1; c >= 0; --c) {
        b[c] /= A[c + (size_t)c * p];
        for (int r = 0; r < c; ++r) b[r] -= A[r + (size_t)c * p] * b[c];
    }
}

// symmetric eigen-decomposition by cyclic Jacobi; eigenvalues in w, eigenvectors in the columns of V
void jacobi_eigh(std::vector<double>& A, int k, std::vector<double>& w, std::vector<double>& V) {
    V.assign((size_t)k * k, 0.0);
    for (int i = 0; i < k; ++i) V[i + (size_t)i * k] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int j = 0; j < k; ++j)
            for (int i = 0; i < k; ++i) (i == j ? diag : off) += A[i + (size_t)j * k] * A[i + (size_t)j * k];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int pI = 0; pI < k - 1; ++pI)
            for (int q = pI + 1; q < k; ++q) {
                const double apq = A[pI + (size_t)q * k];
                if (apq == 0.0) continue;
                const double app = A[pI + (size_t)pI * k], aqq = A[q + (size_t)q * k];
                const double theta = (aqq - app) / (2.0 * apq);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(tt * tt + 1.0), s = tt * c;
                for (int r = 0; r < k; ++r) {   // columns p, q
                    const double arp = A[r + (size_t)pI * k], arq = A[r + (size_t)q * k];
                    A[r + (size_t)pI * k] = c * arp - s * arq;
                    A[r + (size_t)q * k] = s * arp + c * arq;
                }
                for (int r = 0; r < k; ++r) {   // rows p, q
                    const double apr = A[pI + (size_t)r * k], aqr = A[q + (size_t)r * k];
                    A[pI + (size_t)r * k] = c * apr - s * aqr;
                    A[q + (size_t)r * k] = s * apr + c * aqr;
                }
                for (int r = 0; r < k; ++r) {
                    const double vrp = V[r + (size_t)pI * k], vrq = V[r + (size_t)q * k];
                    V[r + (size_t)pI * k] = c * vrp - s * vrq;
                    V[r + (size_t)q * k] = s * vrp + c * vrq;
                }
            }
    }
    w.resize(k);
    for (int i = 0; i < k; ++i) w[i] = A[i + (size_t)i * k];
}

}  // namespace

extern "C" int fgp_bspline_basis(int k, int p, double* B_out) {
    if (k <= 0 || p <= 0 || !B_out) return FGP_ERR_ARG;
    std::vector<double> t;
    int ord;
    bspline_knots(k, p, t, ord);
    std::fill(B_out, B_out + (size_t)k * p, 0.0);
    std::vector<double> N(ord);
    for (int xi = 1; xi <= k; ++xi) {
        const double x = (double)xi;
        // span: last j in [ord-1, p-1] with t[j] <= x  (right end point belongs to the last interval, outer.ok)
        int span = ord - 1;
        while (span < p - 1 && t[span + 1] <= x) ++span;
        basis_funs(t, ord, span, x, N.data());
        for (int j = 0; j < ord; ++j) {
            const int col = span - (ord - 1) + j;
            if (col >= 0 && col < p) B_out[(xi - 1) + (size_t)col * k] = N[j];
        }
    }
    return 0;
}

extern "C" int fgp_project(fgp_ctx* ctx, const double* f, int n, int k, int p, int basis_type, const double* B_in,
                           double* B_out, double* J_out, double* coefs_out) {
    (void)ctx;
    if (!f || n <= 0 || k <= 0 || p < 0 || p > k || !coefs_out) return FGP_ERR_ARG;
    if (p == 0) {   // R/7_dimRedFunctions.R:22-25
        if (B_out) { std::fill(B_out, B_out + (size_t)k * k, 0.0); for (int i = 0; i < k; ++i) B_out[i + (size_t)i * k] = 1.0; }
        if (J_out) { std::fill(J_out, J_out + (size_t)k * k, 0.0); for (int i = 0; i < k; ++i) J_out[i + (size_t)i * k] = 1.0; }
        memcpy(coefs_out, f, sizeof(double) * (size_t)n * k);
        return 0;
    }
    std::vector<double> B((size_t)k * p);
    if (B_in) memcpy(B.data(), B_in, sizeof(double) * (size_t)k * p);
    else if (basis_type == FGP_BASIS_BSPLINES) {
        int rc = fgp_bspline_basis(k, p, B.data());
        if (rc) return rc;
    } else if (basis_type == FGP_BASIS_PCA) {
        if (n < 2) return FGP_ERR_ARG;
        // cov(f): (n-1) denominator  (R/7_dimRedFunctions.R:58)
        std::vector<double> mu(k, 0.0), C((size_t)k * k, 0.0), w, V;
        for (int c = 0; c < k; ++c) {
            double s = 0.0;
            for (int r = 0; r < n; ++r) s += f[r + (size_t)c * n];
            mu[c] = s / n;
        }
        for (int a = 0; a < k; ++a)
            for (int b = a; b < k; ++b) {
                double s = 0.0;
                for (int r = 0; r < n; ++r) s += (f[r + (size_t)a * n] - mu[a]) * (f[r + (size_t)b * n] - mu[b]);
                C[a + (size_t)b * k] = C[b + (size_t)a * k] = s / (n - 1);
            }
        jacobi_eigh(C, k, w, V);
        std::vector<int> ord(k);
        for (int i = 0; i < k; ++i) ord[i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return w[a] > w[b]; });
        for (int j = 0; j < p; ++j) {
            const double* v = &V[(size_t)ord[j] * k];
            // sign convention (LAPACK's is arbitrary, quirk Q7): largest-magnitude component positive
            int im = 0;
            for (int i = 1; i < k; ++i) if (std::fabs(v[i]) > std::fabs(v[im])) im = i;
            const double sg = v[im] < 0 ? -1.0 : 1.0;
            for (int i = 0; i < k; ++i) B[i + (size_t)j * k] = sg * v[i];
        }
    } else return FGP_ERR_ARG;
    // J = B'B ; coefs = t(solve(J, B' f'))   (R/7_dimRedFunctions.R:18-19)
    std::vector<double> Jm((size_t)p * p, 0.0);
    for (int a = 0; a < p; ++a)
        for (int b = 0; b < p; ++b) {
            double s = 0.0;
            for (int i = 0; i < k; ++i) s += B[i + (size_t)a * k] * B[i + (size_t)b * k];
            Jm[a + (size_t)b * p] = s;
        }
    if (B_out) memcpy(B_out, B.data(), sizeof(double) * (size_t)k * p);
    if (J_out) memcpy(J_out, Jm.data(), sizeof(double) * (size_t)p * p);
    std::vector<double> LU = Jm;
    std::vector<int> piv;
    if (!lu_factor(LU, p, piv)) return 1;   // singular Gram matrix: base::solve would stop
    std::vector<double> rhs(p);
    for (int r = 0; r < n; ++r) {
        for (int a = 0; a < p; ++a) {
            double s = 0.0;
            for (int i = 0; i < k; ++i) s += B[i + (size_t)a * k] * f[r + (size_t)i * n];
            rhs[a] = s;
        }
        lu_solve(LU, p, piv, rhs.data());
        for (int a = 0; a < p; ++a) coefs_out[r + (size_t)a * n] = rhs[a];
    }
    return 0;
}
