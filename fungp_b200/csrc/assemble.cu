// Fused distance + correlation assembly.  One pass from point coordinates to covariance entries: the per-input
// distance matrices of the reference (setDistMatrix_S / setDistMatrix_F, L2_byindex, L2_bygroup:
// R/7_distanceFunctions.R:3-46) are never materialised, and the product of per-input correlation factors
// (setR, gaussian_cor / matern52_cor / matern32_cor: R/7_correlFunctions.R:1-80; hybrid product
// R/3_training_SF.R:248; "+ diag(nugget)" :248; "sig2 * (...)" R/4_prediction_SF.R:6,26) is evaluated with a
// single exp per entry:
//   gauss      exp(-1/2 sum_l (d_l/theta_l)^2)
//   matern5_2  prod_l (1 + a_l + a_l^2/3) * exp(-sum_l a_l),  a_l = sqrt5 d_l / theta_l
//   matern3_2  prod_l (1 + a_l)           * exp(-sum_l a_l),  a_l = sqrt3 d_l / theta_l
// Differences are formed first and scaled afterwards (x_i - x_j is exact for close points; scaling the
// coordinates first would lose that), and L2_bygroup keeps the reference's bilinear form
// sum_a (c_ia - c_ja)(e_ia - e_ja), e = c J, clamped at 0 before the square root (R/7_distanceFunctions.R:42-45).
//
// Tiling: 64 x 64 output tile per CTA, 256 threads, 4 x 4 entries per thread arranged so that every store
// instruction writes 256 contiguous bytes per half-warp of the column-major destination; only tiles on or
// below the diagonal are produced for symmetric blocks; coordinates are staged through shared memory in
// chunks of 16 columns.  The kernel is bound by FP64 issue (about 3 DFMA-class ops per unit and entry plus one
// exp), not by HBM -- see DESIGN.md for the op count against the 8 bytes written per entry.
#include "common.cuh"

namespace {

constexpr int AT = 64;       // tile edge
constexpr int CC = 16;       // coordinate columns staged per chunk
constexpr int ATHREADS = 256;

enum { MODE_STORE = 0, MODE_MAX = 1 };

struct AsmK {
    AsmArgs a;
    int u0, u1;            // unit range
    double* maxOut;        // MODE_MAX
};

__device__ __forceinline__ void tri_decode(int t, int& I, int& J) {
    int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    I = i;
    J = t - i * (i + 1) / 2;
}

// KER: 0 = raw distance of the (single) term, 1 gauss, 2 matern5_2, 3 matern3_2
template <int KER, int MODE>
__global__ void __launch_bounds__(ATHREADS) assemble_kernel(AsmK k) {
    __shared__ double sP[CC][AT];
    __shared__ double sQ[CC][AT];
    __shared__ double sRed[ATHREADS / 32];
    const AsmArgs& a = k.a;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    int I, J;
    if (a.sym) tri_decode(blockIdx.x, I, J);
    else {
        const int nTP = (a.nP + AT - 1) / AT;
        I = blockIdx.x % nTP;
        J = blockIdx.x / nTP;
    }
    const int i0 = I * AT, j0 = J * AT;
    const double* scale = a.scale ? a.scale + (long)blockIdx.y * a.nterms : nullptr;

    // this thread's rows (local): 2tx, 2tx+1, 32+2tx, 33+2tx ; cols: 4ty .. 4ty+3
    int lr[4] = {2 * tx, 2 * tx + 1, 32 + 2 * tx, 33 + 2 * tx};
    const int lc0 = 4 * ty;

    double S[4][4], Pr[4][4], D2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { S[i][j] = 0.0; Pr[i][j] = 1.0; D2[i][j] = 0.0; }

    int u = k.u0;
    const int colBegin = (k.u0 < k.u1) ? (a.units[k.u0].col / CC) * CC : 0;
    const int colEnd = (k.u0 < k.u1) ? a.units[k.u1 - 1].col + (a.units[k.u1 - 1].kind ? 2 : 1) : 0;
    for (int cbase = colBegin; cbase < colEnd; cbase += CC) {
        __syncthreads();
        // stage CC columns x 64 points of each side (coalesced along points)
        for (int idx = tid; idx < CC * AT; idx += ATHREADS) {
            const int c = idx / AT, pnt = idx % AT;
            const int col = cbase + c;
            double vp = 0.0, vq = 0.0;
            if (col < colEnd) {
                if (i0 + pnt < a.nP) vp = a.XP[(long)col * a.ldP + i0 + pnt];
                if (j0 + pnt < a.nQ) vq = a.XQ[(long)col * a.ldQ + j0 + pnt];
            }
            sP[c][pnt] = vp;
            sQ[c][pnt] = vq;
        }
        __syncthreads();
        while (u < k.u1) {
            const FgpUnit un = a.units[u];
            if (un.col >= cbase + CC) break;
            const int c = un.col - cbase;
            double xp[4], xq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xp[i] = sP[c][lr[i]];
#pragma unroll
            for (int j = 0; j < 4; ++j) xq[j] = sQ[c][lc0 + j];
            if (un.kind == 0) {
                const double sc = (KER == 0) ? 1.0 : scale[un.term];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double d = fabs(xp[i] - xq[j]);
                        if (KER == 0) S[i][j] = d;
                        else {
                            const double av = d * sc;
                            if (KER == 1) S[i][j] = fma(av, av, S[i][j]);
                            else {
                                S[i][j] += av;
                                if (KER == 2) Pr[i][j] *= fma(av, fma(av, 1.0 / 3.0, 1.0), 1.0);
                                else Pr[i][j] *= (1.0 + av);
                            }
                        }
                    }
            } else {
                double ep[4], eq[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) ep[i] = sP[c + 1][lr[i]];
#pragma unroll
                for (int j = 0; j < 4; ++j) eq[j] = sQ[c + 1][lc0 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) D2[i][j] = fma(xp[i] - xq[j], ep[i] - eq[j], D2[i][j]);
                if (un.last) {
                    const double sc = (KER == 0) ? 1.0 : scale[un.term];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double d2 = fmax(D2[i][j], 0.0);
                            D2[i][j] = 0.0;
                            if (KER == 0) S[i][j] = sqrt(d2);
                            else if (KER == 1) S[i][j] = fma(d2, sc * sc, S[i][j]);
                            else {
                                const double av = sqrt(d2) * sc;
                                S[i][j] += av;
                                if (KER == 2) Pr[i][j] *= fma(av, fma(av, 1.0 / 3.0, 1.0), 1.0);
                                else Pr[i][j] *= (1.0 + av);
                            }
                        }
                }
            }
            ++u;
        }
    }

    if (MODE == MODE_MAX) {
        double mx = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i0 + lr[i] < a.nP && j0 + lc0 + j < a.nQ) mx = fmax(mx, S[i][j]);
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((tid & 31) == 0) sRed[tid >> 5] = mx;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < ATHREADS / 32; ++w) mx = fmax(mx, sRed[w]);
            // non-negative doubles order like their bit patterns
            atomicMax((unsigned long long*)k.maxOut, (unsigned long long)__double_as_longlong(mx));
        }
        return;
    }

    // ---- finish + store
    double* M = a.M + (long)blockIdx.y * a.strideM;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int gj = j0 + lc0 + j;
        if (gj >= a.nQ) continue;
        double v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gi = i0 + lr[i];
            double r;
            if (KER == 0) r = S[i][j];
            else if (KER == 1) r = exp(-0.5 * S[i][j]);
            else r = Pr[i][j] * exp(-S[i][j]);
            const bool dg = a.sym && (gi == gj);
            if (KER != 0) r = a.mult * (r + (dg ? a.diag_rel : 0.0)) + (dg ? a.diag_abs : 0.0);
            v[i] = r;
        }
        double* colp = M + (long)(a.colOff + gj) * a.ld + a.rowOff;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int gi = i0 + lr[2 * h];
            if (gi + 1 < a.nP && ((reinterpret_cast<uintptr_t>(colp + gi) & 15) == 0)) {
                *reinterpret_cast<double2*>(colp + gi) = make_double2(v[2 * h], v[2 * h + 1]);
            } else {
                if (gi < a.nP) colp[gi] = v[2 * h];
                if (gi + 1 < a.nP) colp[gi + 1] = v[2 * h + 1];
            }
        }
        if (a.sym && a.mirror) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int gi = i0 + lr[i];
                if (gi < a.nP) M[(long)(a.colOff + gi) * a.ld + a.rowOff + gj] = v[i];
            }
        }
    }
}

template <int MODE>
void dispatch(const AsmK& k, int ker, dim3 grid, cudaStream_t st) {
    switch (ker) {
        case 0: FGP_COUNT_LAUNCH(); assemble_kernel<0, MODE><<<grid, ATHREADS, 0, st>>>(k); break;
        case 1: FGP_COUNT_LAUNCH(); assemble_kernel<1, MODE><<<grid, ATHREADS, 0, st>>>(k); break;
        case 2: FGP_COUNT_LAUNCH(); assemble_kernel<2, MODE><<<grid, ATHREADS, 0, st>>>(k); break;
        default: FGP_COUNT_LAUNCH(); assemble_kernel<3, MODE><<<grid, ATHREADS, 0, st>>>(k); break;
    }
}

int ntiles_of(const AsmArgs& a) {
    const int nTP = (a.nP + AT - 1) / AT, nTQ = (a.nQ + AT - 1) / AT;
    return a.sym ? nTP * (nTP + 1) / 2 : nTP * nTQ;
}

}  // namespace

void launch_assemble(const AsmArgs& a, int batch, cudaStream_t st) {
    if (a.nP <= 0 || a.nQ <= 0 || batch <= 0) return;
    AsmK k{a, 0, a.nunits, nullptr};
    dispatch<MODE_STORE>(k, a.ker, dim3(ntiles_of(a), batch), st);
}

void launch_distmax(const double* X, int n, long ldX, const FgpUnit* units, int u0, int u1, double* out_max,
                    cudaStream_t st) {
    AsmArgs a{};
    a.XP = X; a.nP = n; a.ldP = ldX; a.XQ = X; a.nQ = n; a.ldQ = ldX;
    a.units = units; a.nunits = u1; a.sym = 1; a.ker = 0;
    AsmK k{a, u0, u1, out_max};
    dispatch<MODE_MAX>(k, 0, dim3(ntiles_of(a), 1), st);
}

void launch_distmat(const double* X, int n, long ldX, const FgpUnit* units, int u0, int u1, double* out, long ld,
                    cudaStream_t st) {
    AsmArgs a{};
    a.XP = X; a.nP = n; a.ldP = ldX; a.XQ = X; a.nQ = n; a.ldQ = ldX;
    a.units = units; a.nunits = u1; a.sym = 0; a.ker = 0;
    a.M = out; a.ld = ld; a.strideM = 0; a.rowOff = 0; a.colOff = 0;
    AsmK k{a, u0, u1, nullptr};
    dispatch<MODE_STORE>(k, 0, dim3(ntiles_of(a), 1), st);
}
