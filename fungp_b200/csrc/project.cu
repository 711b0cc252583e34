// Dimension reduction of functional inputs: B-spline design matrix, PCA basis and the projection
// coefs = t(solve(B'B, B' f')).  Follows R/7_dimRedFunctions.R:7-60; the arithmetic the reference delegates to
// splines::splineDesign, stats::cov, base::eigen and base::solve is restated here (Cox-de Boor recursion, cyclic
// Jacobi, LU with partial pivoting).
// With a context the two contractions over the n curves run on the device -- the covariance of proj_pca
// (R/7_dimRedFunctions.R:58) and the projection itself, written as coefs = f * (B J^-1) (:18-19) -- while the k x p
// design matrix, the k x k eigen-decomposition and the p x p solve stay on the host (k <= a few hundred).  fgp_project
// requires a context; the context-free loops survive only behind the test-support symbol fgp_test_project_host.
#include "../../include/fungp_b200.h"
#include "ctx.h"
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

// ---- device contractions ------------------------------------------------------------------------------------------
// column means of the n x k curve matrix (one block per column, fixed reduction order)
__global__ void __launch_bounds__(256) proj_colmean_kernel(const double* __restrict__ f, int n, int k,
                                                           double* __restrict__ mu) {
    __shared__ double red[256];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int r = threadIdx.x; r < n; r += 256) s += f[r + (size_t)c * n];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) mu[c] = red[0] / n;
}

// C = (f - mu)'(f - mu) / (n - 1), 4 x 4 entries per block on and above the diagonal, mirrored on store
__global__ void __launch_bounds__(256) proj_cov_kernel(const double* __restrict__ f, const double* __restrict__ mu,
                                                       int n, int k, double* __restrict__ C) {
    const int a0 = blockIdx.x * 4, b0 = blockIdx.y * 4;
    if (b0 + 3 < a0) return;
    __shared__ double red[8][16];
    double ma[4], mb[4], acc[16];
    for (int j = 0; j < 4; ++j) {
        ma[j] = a0 + j < k ? mu[a0 + j] : 0.0;
        mb[j] = b0 + j < k ? mu[b0 + j] : 0.0;
    }
    for (int j = 0; j < 16; ++j) acc[j] = 0.0;
    for (int r = threadIdx.x; r < n; r += 256) {
        double va[4], vb[4];
        for (int j = 0; j < 4; ++j) {
            va[j] = a0 + j < k ? f[r + (size_t)(a0 + j) * n] - ma[j] : 0.0;
            vb[j] = b0 + j < k ? f[r + (size_t)(b0 + j) * n] - mb[j] : 0.0;
        }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fma(va[i], vb[j], acc[i * 4 + j]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = 0; j < 16; ++j) {
        double v = acc[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        const int a = a0 + (threadIdx.x >> 2), b = b0 + (threadIdx.x & 3);
        if (a < k && b < k && a <= b) {
            v /= (double)(n - 1);
            C[a + (size_t)b * k] = v;
            C[b + (size_t)a * k] = v;
        }
    }
}

// coefs (n x p) = f (n x k) * M (k x p): one curve per thread, 8 coefficients per thread
__global__ void __launch_bounds__(128) proj_apply_kernel(const double* __restrict__ f, const double* __restrict__ M,
                                                         int n, int k, int p, double* __restrict__ coefs) {
    const int r = blockIdx.x * 128 + threadIdx.x, a0 = blockIdx.y * 8;
    if (r >= n) return;
    double acc[8];
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    const int na = min(8, p - a0);
    for (int i = 0; i < k; ++i) {
        const double fv = f[r + (size_t)i * n];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < na) acc[j] = fma(fv, M[i + (size_t)(a0 + j) * k], acc[j]);
    }
    for (int j = 0; j < na; ++j) coefs[r + (size_t)(a0 + j) * n] = acc[j];
}

#define PCK(expr)                                                             \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            fgp_set_err_cuda(ctx, #expr, _e, __FILE__, __LINE__);             \
            return FGP_ERR_CUDA;                                              \
        }                                                                     \
    } while (0)

// blocks of the context's pool, handed back when the call leaves
struct PoolBlocks {
    DevPool& pool;
    std::vector<void*> held;
    explicit PoolBlocks(DevPool& p) : pool(p) {}
    double* get(size_t count) {
        void* q = pool.get(sizeof(double) * count);
        if (q) held.push_back(q);
        return (double*)q;
    }
    ~PoolBlocks() { for (void* q : held) pool.put(q); }
};

int device_covariance(fgp_ctx* ctx, const double* df, int n, int k, PoolBlocks& blk, std::vector<double>& C) {
    DevState& d = ctx->dev[0];
    cudaStream_t st = d.sPanel[0];
    double* dmu = blk.get(k);
    double* dC = blk.get((size_t)k * k);
    if (!dmu || !dC) { fgp_set_err(ctx, "out of device memory"); return FGP_ERR_ALLOC; }
    proj_colmean_kernel<<<k, 256, 0, st>>>(df, n, k, dmu);
    FGP_COUNT_LAUNCH();
    const int t = (k + 3) / 4;
    proj_cov_kernel<<<dim3(t, t), 256, 0, st>>>(df, dmu, n, k, dC);
    FGP_COUNT_LAUNCH();
    PCK(cudaGetLastError());
    C.resize((size_t)k * k);
    PCK(cudaMemcpyAsync(C.data(), dC, sizeof(double) * (size_t)k * k, cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    return 0;
}

int device_apply(fgp_ctx* ctx, const double* df, const std::vector<double>& M, int n, int k, int p, PoolBlocks& blk,
                 double* coefs_out) {
    DevState& d = ctx->dev[0];
    cudaStream_t st = d.sPanel[0];
    double* dM = blk.get((size_t)k * p);
    double* dc = blk.get((size_t)n * p);
    if (!dM || !dc) { fgp_set_err(ctx, "out of device memory"); return FGP_ERR_ALLOC; }
    PCK(cudaMemcpyAsync(dM, M.data(), sizeof(double) * (size_t)k * p, cudaMemcpyHostToDevice, st));
    proj_apply_kernel<<<dim3((n + 127) / 128, (p + 7) / 8), 128, 0, st>>>(df, dM, n, k, p, dc);
    FGP_COUNT_LAUNCH();
    PCK(cudaGetLastError());
    PCK(cudaMemcpyAsync(coefs_out, dc, sizeof(double) * (size_t)n * p, cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    return 0;
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

static int project_impl(fgp_ctx* ctx, const double* f, int n, int k, int p, int basis_type, const double* B_in,
                        double* B_out, double* J_out, double* coefs_out) {
    if (!f || n <= 0 || k <= 0 || p < 0 || p > k || !coefs_out) return FGP_ERR_ARG;
    if (p == 0) {   // R/7_dimRedFunctions.R:22-25
        if (B_out) { std::fill(B_out, B_out + (size_t)k * k, 0.0); for (int i = 0; i < k; ++i) B_out[i + (size_t)i * k] = 1.0; }
        if (J_out) { std::fill(J_out, J_out + (size_t)k * k, 0.0); for (int i = 0; i < k; ++i) J_out[i + (size_t)i * k] = 1.0; }
        memcpy(coefs_out, f, sizeof(double) * (size_t)n * k);
        return 0;
    }
    // device copy of the curves (only when a context is given)
    std::unique_ptr<PoolBlocks> blk;
    double* df = nullptr;
    if (ctx) {
        if (ctx->dev.empty()) return FGP_ERR_NODEVICE;
        DevState& d = ctx->dev[0];
        PCK(cudaSetDevice(d.device));
        blk.reset(new PoolBlocks(d.pool));
        df = blk->get((size_t)n * k);
        if (!df) { fgp_set_err(ctx, "out of device memory"); return FGP_ERR_ALLOC; }
        PCK(cudaMemcpyAsync(df, f, sizeof(double) * (size_t)n * k, cudaMemcpyHostToDevice, d.sPanel[0]));
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
        if (ctx) {
            int rc = device_covariance(ctx, df, n, k, *blk, C);
            if (rc) return rc;
        } else {
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
    if (ctx) {
        // M = B J^-1 (k x p; J is symmetric, so row i of M is J^-1 applied to row i of B), then coefs = f M
        std::vector<double> M((size_t)k * p);
        for (int i = 0; i < k; ++i) {
            for (int a = 0; a < p; ++a) rhs[a] = B[i + (size_t)a * k];
            lu_solve(LU, p, piv, rhs.data());
            for (int a = 0; a < p; ++a) M[i + (size_t)a * k] = rhs[a];
        }
        return device_apply(ctx, df, M, n, k, p, *blk, coefs_out);
    }
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

// The product entry point always runs the contractions over the curves on the device: a context is required.
extern "C" int fgp_project(fgp_ctx* ctx, const double* f, int n, int k, int p, int basis_type, const double* B_in,
                           double* B_out, double* J_out, double* coefs_out) {
    if (!ctx) return FGP_ERR_ARG;
    return project_impl(ctx, f, n, k, p, basis_type, B_in, B_out, J_out, coefs_out);
}

// TEST SUPPORT ONLY (tests/test_cabi_and_host.py, runs without a GPU): the host-side pieces of dimReduction -- B-spline
// design matrix, Jacobi eigenvectors, LU solve -- with the two contractions as plain loops, so that they can be checked
// against the oracle and the reference's stored bases on a machine that has no device.  Nothing in the library calls it.
extern "C" int fgp_test_project_host(const double* f, int n, int k, int p, int basis_type, const double* B_in,
                                     double* B_out, double* J_out, double* coefs_out) {
    return project_impl(nullptr, f, n, k, p, basis_type, B_in, B_out, J_out, coefs_out);
}
