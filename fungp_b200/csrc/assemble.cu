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
// coordinates first would lose that).  L2_bygroup terms arrive whitened (z = c chol(J), api.cu build_coords), so
// that the reference's bilinear form (c_i - c_j)' J (c_i - c_j) (R/7_distanceFunctions.R:42-45) is a plain sum of
// squares -- one coordinate column, one subtraction and one fused multiply-add per coefficient and entry.
//
// Tiling: 64 x 64 output tile per CTA, 256 threads, 4 x 4 entries per thread arranged so that every store
// instruction writes 256 contiguous bytes per half-warp of the column-major destination; only tiles on or
// below the diagonal are produced for symmetric blocks; up to 32 coordinate columns of both point sets and the
// metadata of their units are staged through shared memory per pass (one pass for every BASELINE configuration),
// read back with 16-byte loads.  The kernel is bound by FP64 issue, not by HBM: about 2-3 FP64 instructions per
// coordinate and entry plus an 18-instruction exp against 8 bytes written, on a machine whose DFMA and DMMA share
// one FP64 datapath (profiles/r01_peaks.md) -- see DESIGN.md 2.3 for the count.
#include "common.cuh"
#include <atomic>
#include <algorithm>

namespace {

constexpr int AT = 64;       // tile edge
constexpr int CC = 32;       // coordinate columns staged per pass
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

// exp(t) for t <= 0, table-driven and branch-free:  t = (32 e + j) ln2/32 + r,  |r| <= ln2/64,
//   exp(t) = 2^e * 2^(j/32) * exp(r),   2^(j/32) = T_hi[j] + T_lo[j] from a 32-entry table (shared memory, one 16-byte load),
//   exp(r) - 1 = r p(r), p of degree 5 (truncation r^7/5040 < 4e-18 relative).
// 12 FP64 instructions and a dependent chain of 9 instead of 18 / 16 for the degree-12 polynomial it replaces, and a
// smaller error: <= 0.9 ulp against the exact value in a numpy emulation over 3 M arguments in [-700, 0]
// (tests/test_gpu_parity.py::test_assembly_exp_accuracy checks the kernel against numpy's exp at 2 ulp).  Below -708 the
// result is flushed to 0 (the true value is < 3.4e-308).
__device__ const double2 g_exp_tab[32] = {
    {1.00000000000000000e+00, 0.00000000000000000e+00},
    {1.02189714865411663e+00, 5.10922502897344389e-17},
    {1.04427378242741375e+00, 8.55188970553796489e-17},
    {1.06714040067682370e+00, -7.89985396684158212e-17},
    {1.09050773266525769e+00, -3.04678207981247115e-17},
    {1.11438674259589243e+00, 1.04102784568455710e-16},
    {1.13878863475669156e+00, 8.91281267602540778e-17},
    {1.16372485877757748e+00, 3.82920483692409350e-17},
    {1.18920711500272103e+00, 3.98201523146564611e-17},
    {1.21524735998046896e+00, -7.71263069268148813e-17},
    {1.24185781207348400e+00, 4.65802759183693679e-17},
    {1.26905095719173322e+00, 2.66793213134218610e-18},
    {1.29683955465100964e+00, 2.53825027948883150e-17},
    {1.32523664315974132e+00, -2.85873121003886137e-17},
    {1.35425554693689265e+00, 7.70094837980298946e-17},
    {1.38390988196383202e+00, -6.77051165879478629e-17},
    {1.41421356237309515e+00, -9.66729331345291345e-17},
    {1.44518080697704665e+00, -3.02375813499398732e-17},
    {1.47682614593949935e+00, -3.48399455689279580e-17},
    {1.50916442759342284e+00, -1.01645532775429504e-16},
    {1.54221082540794074e+00, 7.94983480969762086e-17},
    {1.57598084510788650e+00, -1.01369164712783040e-17},
    {1.61049033194925428e+00, 2.47071925697978879e-17},
    {1.64575547815396495e+00, -1.01256799136747726e-16},
    {1.68179283050742900e+00, 8.19901002058149652e-17},
    {1.71861929812247793e+00, -1.85138041826311099e-17},
    {1.75625216037329945e+00, 2.96014069544887331e-17},
    {1.79470907500310717e+00, 1.82274584279120868e-17},
    {1.83400808640934243e+00, 3.28310722424562720e-17},
    {1.87416763411029996e+00, -6.12276341300414256e-17},
    {1.91520656139714740e+00, -1.06199460561959626e-16},
    {1.95714412417540018e+00, 8.96076779103666777e-17}};

// The table sits in shared memory eight times over, entry e of copy c at 16-byte slot 8 e + c: a lane only ever reads
// copy (lane & 7), so the eight lanes of a quarter-warp (the unit a 16-byte shared load is served in) hit eight
// different bank groups whatever their arguments are.  With a single copy the data-dependent indices of neighbouring
// entries collided: ncu counted 96 M bank-conflict cycles per 436 M lookups, ~7 extra wavefronts per warp load.
constexpr int EXP_COPIES = 8;
constexpr int EXP_TAB_SLOTS = 32 * EXP_COPIES;
__device__ __forceinline__ void load_exp_table(double2* sTab) {
    if (threadIdx.x < EXP_TAB_SLOTS) sTab[threadIdx.x] = g_exp_tab[threadIdx.x / EXP_COPIES];
}
// this lane's copy: pass the result to exp_nonpos
__device__ __forceinline__ const double2* exp_table_of_lane(const double2* sTab) { return sTab + (threadIdx.x & (EXP_COPIES - 1)); }

__device__ __forceinline__ double exp_nonpos(double t, const double2* sTab) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: rounds to nearest integer in the low word
    const double kd = fma(t, 4.61662413084468284e+01, MAGIC);          // 32 / ln2
    const int k = __double2loint(kd);
    const double kf = kd - MAGIC;
    double r = fma(kf, -2.16608493865351193e-02, t);                   // ln2/32, high part (32 significant bits: kf * hi exact)
    r = fma(kf, -5.96317165397058656e-12, r);                          // low part
    double p = 1.38888888888888888889e-03;                 // 1/6!
    p = fma(p, r, 8.33333333333333333333e-03);             // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);             // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);             // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    const double2 T = sTab[(k & 31) * EXP_COPIES];
    double res = fma(T.x, __dmul_rn(r, p), T.x);           // T_hi (1 + r p)
    res = __dadd_rn(res, T.y);
    // below about -708.0002 the result would leave the normal range (the exponent is patched by integer addition): flush to
    // zero.  Negative doubles order like their high words taken as unsigned integers; 0xC0862000 is the high word of -708.
    const bool under = (unsigned)__double2hiint(t) > 0xC0862000u;
    const int hi = __double2hiint(res) + ((k >> 5) << 20);
    return __hiloint2double(under ? 0 : hi, under ? 0 : __double2loint(res));
}

// One length-scale term of a Matern kernel, a = d * sc (sc = sqrt(3 or 5) / theta): the argument of the single exp
// gathers -a, the polynomial factors multiply up.  matern3_2: 1 + a;  matern5_2: 1 + a + a^2 / 3 = 1 + d (sc + d c2),
// c2 = sc^2 / 3.  Three / four FP64 operations per term; both assembly kernels call this, so they round alike.
__device__ __forceinline__ double matern52_c2(double sc) { return __dmul_rn(__dmul_rn(sc, sc), 1.0 / 3.0); }
template <int KER>
__device__ __forceinline__ void matern_term(double d, double sc, double c2, double& S, double& Pr) {
    S = fma(d, -sc, S);
    if (KER == 2) Pr = __dmul_rn(Pr, fma(d, fma(d, c2, sc), 1.0));
    else Pr = __dmul_rn(Pr, fma(d, sc, 1.0));
}

// KER: 0 = raw distance of the (single) term, 1 gauss, 2 matern5_2, 3 matern3_2
template <int KER, int MODE>
__global__ void __launch_bounds__(ATHREADS, 2) assemble_kernel(AsmK k) {
    __shared__ __align__(16) double sP[CC][AT];
    __shared__ __align__(16) double sQ[CC][AT];
    __shared__ double sScale[CC], sScale2[CC];
    __shared__ int sCol[CC];
    __shared__ int sFlag[CC];       // bit 0: member of a group term, bit 1: last unit of its term
    __shared__ double sRed[ATHREADS / 32];
    __shared__ double2 sExpAll[EXP_TAB_SLOTS];
    if (KER != 0) load_exp_table(sExpAll);   // visible after the barriers of the staging loop below
    const double2* sExp = exp_table_of_lane(sExpAll);
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
    const double* XP = a.XP;
    const double* XQ = a.XQ;
    const FgpUnit* units = a.units;
    int u0 = k.u0, u1 = k.u1;
    if (a.elems) {
        // heterogeneous batch: this element's own structure (the points of many candidate fits share one launch)
        const AsmElem e = a.elems[blockIdx.y];
        XP = XQ = e.X; units = e.units; scale = e.scale; u0 = 0; u1 = e.nunits;
    }

    // this thread's rows (local): 2tx, 2tx+1, 32+2tx, 33+2tx ; cols: 4ty .. 4ty+3
    const int lr[4] = {2 * tx, 2 * tx + 1, 32 + 2 * tx, 33 + 2 * tx};
    const int lc0 = 4 * ty;

    double S[4][4], Pr[4][4], D2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { S[i][j] = 0.0; Pr[i][j] = 1.0; D2[i][j] = 0.0; }

    for (int ub = u0; ub < u1; ub += CC) {
        const int nc = min(CC, u1 - ub);
        __syncthreads();
        if (tid < nc) {
            const FgpUnit un = units[ub + tid];
            sCol[tid] = un.col;
            sFlag[tid] = (un.kind ? 1 : 0) | (un.last ? 2 : 0);
            const double sc = (KER == 0) ? 1.0 : scale[un.term];
            // gauss: -theta^-2 / 2 multiplies the SQUARED difference, so the sum arrives as the argument of the exp
            sScale[tid] = (KER == 1) ? __dmul_rn(-0.5, __dmul_rn(sc, sc)) : sc;
            if (KER == 2) sScale2[tid] = matern52_c2(sc);
        }
        __syncthreads();
        // stage nc columns x 64 points of each side (coalesced along points)
        for (int idx = tid; idx < nc * AT; idx += ATHREADS) {
            const int c = idx / AT, pnt = idx % AT;
            const long col = sCol[c];
            sP[c][pnt] = (i0 + pnt < a.nP) ? XP[col * a.ldP + i0 + pnt] : 0.0;
            sQ[c][pnt] = (j0 + pnt < a.nQ) ? XQ[col * a.ldQ + j0 + pnt] : 0.0;
        }
        __syncthreads();
#pragma unroll 2
        for (int c = 0; c < nc; ++c) {
            const double2 p01 = *reinterpret_cast<const double2*>(&sP[c][2 * tx]);
            const double2 p23 = *reinterpret_cast<const double2*>(&sP[c][32 + 2 * tx]);
            const double2 q01 = *reinterpret_cast<const double2*>(&sQ[c][lc0]);
            const double2 q23 = *reinterpret_cast<const double2*>(&sQ[c][lc0 + 2]);
            const double xp[4] = {p01.x, p01.y, p23.x, p23.y};
            const double xq[4] = {q01.x, q01.y, q23.x, q23.y};
            const int flag = sFlag[c];
            const double sc = sScale[c];
            const double sc2 = (KER == 2) ? sScale2[c] : 0.0;
            if (!(flag & 1)) {
                // index unit: scalar input or one L2_byindex coefficient, its own length-scale
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double d = fabs(xp[i] - xq[j]);
                        if (KER == 0) S[i][j] = d;
                        else {
                            // explicit roundings (no contraction left to the compiler): assemble_fused_kernel repeats these
                            // operations one for one and must give the same bits
                            if (KER == 1) S[i][j] = fma(__dmul_rn(d, d), sc, S[i][j]);
                            else matern_term<KER>(d, sc, sc2, S[i][j], Pr[i][j]);
                        }
                    }
            } else {
                // member of an L2_bygroup term (whitened coefficients): squared differences add up
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double d = xp[i] - xq[j];
                        D2[i][j] = fma(d, d, D2[i][j]);
                    }
                if (flag & 2) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double d2 = D2[i][j];
                            D2[i][j] = 0.0;
                            if (KER == 0) S[i][j] = sqrt(d2);
                            else if (KER == 1) S[i][j] = fma(d2, sc, S[i][j]);
                            else matern_term<KER>(sqrt(d2), sc, sc2, S[i][j], Pr[i][j]);
                        }
                }
            }
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
    const bool scaled = a.mult != 1.0;
    const bool diagTile = a.sym && I == J;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int gj = j0 + lc0 + j;
        if (gj >= a.nQ) continue;
        double v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double r;
            if (KER == 0) r = S[i][j];
            else if (KER == 1) r = exp_nonpos(S[i][j], sExp);
            else r = __dmul_rn(Pr[i][j], exp_nonpos(S[i][j], sExp));
            if (KER != 0) {
                if (diagTile && lr[i] == lc0 + j) r = fma(a.mult, __dadd_rn(r, a.diag_rel), a.diag_abs);
                else if (scaled) r = __dmul_rn(r, a.mult);
            }
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

// ---------------------------------------------------------------------------------------------------------------
// Batch-fused variant.  The B points of one finite-difference gradient (theta, theta +- 1e-3 e_i), of a presample or of a
// multistart share every coordinate difference -- only the length-scales differ -- so a tile computes the per-term base
// quantities ONCE (|x_i - x_j| of an index unit; the summed squares, or their root for the Matern kernels, of a group
// term) and then runs B times over (scale, accumulate, exp, store).  Per matrix and entry that leaves 1-2 FP64
// instructions per term + the 18 of the exp instead of 2-3 per coordinate + 18: the kernel moves from the FP64-issue
// ceiling towards the HBM write roofline.  The arithmetic per entry is instruction for instruction that of
// assemble_kernel (same operations in the same order), so both kernels produce bit-identical matrices.
// Tile 64 rows x 16 columns, 256 threads, 4 x 1 entries per thread (the base quantities of 4 entries x up to 8 terms
// live in registers; 124 registers, so two CTAs share an SM and hide each other's staging); structures with more than 8
// length-scales or 48 units, and batches of fewer than 4 points (too little to amortise the per-tile set-up), use
// assemble_kernel.
constexpr int FR = 64, FC = 16;          // fused tile
constexpr int FTHREADS = 256;            // 8 warps: 4 rows x 1 column per thread; two CTAs per SM hide each other's set-up
constexpr int FCC = 48;                  // units (coordinate columns) a fused tile can stage
constexpr int FNT = 8;                   // length-scale terms
constexpr int FBCH = 32;                 // batch elements whose multipliers are staged at a time

struct FusedK {
    AsmArgs a;
    const AsmGroup* groups;   // null: one group = elements [0, count)
    int count;
};

// Persistent: a CTA walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ... of ONE group.  The unit table and -- for
// groups of at most FBCH elements -- the multipliers are staged once per CTA, and the coordinates of the next tile are
// fetched with cp.async into the second half of a double buffer while the B passes over the current tile run: the
// per-tile set-up (three dependent global round trips, measured at ~3.7 passes' worth) is off the critical path.
template <int KER>
__global__ void __launch_bounds__(FTHREADS, 2) assemble_fused_kernel(FusedK k, int ntiles) {
    // double-buffered coordinates in dynamic shared memory (60 KB: above the static limit): sP[2][FCC][FR] | sQ[2][FCC][FC]
    extern __shared__ __align__(16) double fusedDyn[];
    double (*sP)[FCC][FR] = reinterpret_cast<double (*)[FCC][FR]>(fusedDyn);
    double (*sQ)[FCC][FC] = reinterpret_cast<double (*)[FCC][FC]>(fusedDyn + 2 * FCC * FR);
    __shared__ __align__(16) double sSc[FBCH][FNT];   // gauss: -scale^2 / 2 (multiplies squared differences); Matern: scale
    __shared__ __align__(16) double sSc2[KER == 2 ? FBCH : 1][FNT];   // matern5_2: scale^2 / 3
    __shared__ int sCol[FCC], sFlag[FCC], sTermFirst[FNT + 1], sNterms;
    __shared__ double2 sExpAll[EXP_TAB_SLOTS];
    load_exp_table(sExpAll);
    const double2* sExp = exp_table_of_lane(sExpAll);
    const AsmArgs& a = k.a;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;             // rows 4tx..4tx+3, column ty
    constexpr int R = FR / FC;
    AsmGroup g;
    if (k.groups) g = k.groups[blockIdx.y];
    else { g.first = 0; g.count = k.count; }
    const double* X = a.XP;
    const FgpUnit* units = a.units;
    int nunits = a.nunits;
    if (a.elems) { const AsmElem e = a.elems[g.first]; X = e.X; units = e.units; nunits = e.nunits; }
    if (tid < nunits) {
        const FgpUnit un = units[tid];
        sCol[tid] = un.col;
        sFlag[tid] = (un.kind ? 1 : 0) | (un.last ? 2 : 0);
    }
    __syncthreads();
    if (tid < 32) {
        // first unit of every term: a term ends at a unit flagged `last`; prefix counts by ballot (nunits <= FCC <= 64)
        const bool l0 = tid < nunits && (sFlag[tid] & 2), l1 = tid + 32 < nunits && (sFlag[tid + 32] & 2);
        const unsigned m0 = __ballot_sync(0xffffffffu, l0), m1 = __ballot_sync(0xffffffffu, l1);
        const unsigned below = (1u << tid) - 1u;
        if (l0) sTermFirst[__popc(m0 & below) + 1] = tid + 1;
        if (l1) sTermFirst[__popc(m0) + __popc(m1 & below) + 1] = tid + 33;
        if (tid == 0) { sTermFirst[0] = 0; sNterms = __popc(m0) + __popc(m1); }
    }
    // tile t -> (I over 64-row blocks, J over 16-column blocks) with J < R (I + 1):  t = R I (I + 1) / 2 + J
    auto tile_origin = [&](int t, int& i0, int& j0) {
        int I = (int)((sqrtf(8.0f * (float)t / (float)R + 1.0f) - 1.0f) * 0.5f);
        while (R * (I + 1) * (I + 2) / 2 <= t) ++I;
        while (R * I * (I + 1) / 2 > t) --I;
        i0 = I * FR; j0 = (t - R * I * (I + 1) / 2) * FC;
    };
    // coordinates of a tile's 64 + 16 points, every unit: asynchronous 8-byte copies (zeros beyond the last point)
    auto fetch = [&](int t, int buf) {
        int i0, j0;
        tile_origin(t, i0, j0);
        for (int idx = tid; idx < nunits * FR; idx += FTHREADS) {
            const int c = idx / FR, pnt = idx % FR;
            if (i0 + pnt < a.nP) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sP[buf][c][pnt]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(X + (long)sCol[c] * a.ldP + i0 + pnt));
            } else sP[buf][c][pnt] = 0.0;
        }
        for (int idx = tid; idx < nunits * FC; idx += FTHREADS) {
            const int c = idx / FC, pnt = idx % FC;
            if (j0 + pnt < a.nQ) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sQ[buf][c][pnt]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(X + (long)sCol[c] * a.ldQ + j0 + pnt));
            } else sQ[buf][c][pnt] = 0.0;
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    // multipliers of the batch elements [b0, b0 + nb)
    auto stage_scales = [&](int b0, int nb, int nterms) {
        for (int idx = tid; idx < nb * FNT; idx += FTHREADS) {
            const int b = idx / FNT, t = idx % FNT;
            double sc = 0.0;
            if (t < nterms) {
                const double* scale = a.elems ? a.elems[g.first + b0 + b].scale : a.scale + (long)(g.first + b0 + b) * a.nterms;
                sc = scale[t];
            }
            sSc[b][t] = (KER == 1) ? __dmul_rn(-0.5, __dmul_rn(sc, sc)) : sc;
            if (KER == 2) sSc2[b][t] = matern52_c2(sc);
        }
    };
    __syncthreads();                                     // sCol, sNterms
    const int nterms = sNterms;
    const bool scalesOnce = g.count <= FBCH;
    if (scalesOnce) stage_scales(0, g.count, nterms);
    int tile = blockIdx.x;
    if (tile < ntiles) fetch(tile, 0);
    const bool scaled = a.mult != 1.0;

    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();                                 // this tile's coordinates (and, first time, the multipliers) are in; the
                                                         // other buffer is no longer read
        if (tile + (int)gridDim.x < ntiles) fetch(tile + gridDim.x, buf ^ 1);
        int i0, j0;
        tile_origin(tile, i0, j0);

        // ---- base quantity of every term: |x_i - x_j| (squared for gauss), or the summed squares (gauss) / their root
        //      (Matern) of a group
        double q[4][FNT];
#pragma unroll
        for (int t = 0; t < FNT; ++t) {
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i][t] = 0.0;
            if (t < nterms) {
                const int u0 = sTermFirst[t], u1 = sTermFirst[t + 1];
                if (!(sFlag[u0] & 1)) {
                    const double2 p01 = *reinterpret_cast<const double2*>(&sP[buf][u0][4 * tx]);
                    const double2 p23 = *reinterpret_cast<const double2*>(&sP[buf][u0][4 * tx + 2]);
                    const double xq = sQ[buf][u0][ty];
                    q[0][t] = fabs(p01.x - xq); q[1][t] = fabs(p01.y - xq); q[2][t] = fabs(p23.x - xq); q[3][t] = fabs(p23.y - xq);
                    if (KER == 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) q[i][t] = __dmul_rn(q[i][t], q[i][t]);
                    }
                } else {
                    double D2[4] = {0.0, 0.0, 0.0, 0.0};
                    for (int u = u0; u < u1; ++u) {
                        const double2 p01 = *reinterpret_cast<const double2*>(&sP[buf][u][4 * tx]);
                        const double2 p23 = *reinterpret_cast<const double2*>(&sP[buf][u][4 * tx + 2]);
                        const double xq = sQ[buf][u][ty];
                        const double d0 = p01.x - xq, d1 = p01.y - xq, d2 = p23.x - xq, d3 = p23.y - xq;
                        D2[0] = fma(d0, d0, D2[0]); D2[1] = fma(d1, d1, D2[1]); D2[2] = fma(d2, d2, D2[2]); D2[3] = fma(d3, d3, D2[3]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) q[i][t] = (KER == 1) ? D2[i] : sqrt(D2[i]);
                }
            }
        }

        const int gi0 = i0 + 4 * tx, gj = j0 + ty;
        // which of this thread's four entries lies on the diagonal of the matrix (0..3; anything else: none)
        const unsigned dgi = (unsigned)((a.colOff + gj) - (a.rowOff + gi0));
        const long colOffset = (long)(a.colOff + gj) * a.ld + a.rowOff + gi0;
        // 16-byte stores need an aligned address in every matrix of the group: an even stride keeps the alignment of the first
        const bool full4 = gi0 + 3 < a.nP && !(a.strideM & 1) &&
                           ((reinterpret_cast<uintptr_t>(a.M + (long)g.first * a.strideM + colOffset) & 15) == 0);
        for (int b0 = 0; b0 < g.count; b0 += FBCH) {
            const int nb = min(FBCH, g.count - b0);
            if (!scalesOnce) {
                __syncthreads();
                stage_scales(b0, nb, nterms);
                __syncthreads();
            }
            if (gj >= a.nQ) continue;
            double* colp = a.M + (long)(g.first + b0) * a.strideM + colOffset;
            for (int b = 0; b < nb; ++b, colp += a.strideM) {
                double S[4] = {0.0, 0.0, 0.0, 0.0}, Pr[4] = {1.0, 1.0, 1.0, 1.0};
                double scv[FNT];
#pragma unroll
                for (int t = 0; t < FNT; t += 2) {
                    const double2 s2 = *reinterpret_cast<const double2*>(&sSc[b][t]);
                    scv[t] = s2.x; scv[t + 1] = s2.y;
                }
#pragma unroll
                for (int t = 0; t < FNT; ++t) {
                    if (t < nterms) {
                        const double sc = scv[t];
                        if (KER == 1) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) S[i] = fma(q[i][t], sc, S[i]);
                        } else {
                            const double c2 = (KER == 2) ? sSc2[KER == 2 ? b : 0][t] : 0.0;
#pragma unroll
                            for (int i = 0; i < 4; ++i) matern_term<KER>(q[i][t], sc, c2, S[i], Pr[i]);
                        }
                    }
                }
                double raw[4], v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    raw[i] = (KER == 1) ? exp_nonpos(S[i], sExp) : __dmul_rn(Pr[i], exp_nonpos(S[i], sExp));
                    v[i] = scaled ? __dmul_rn(raw[i], a.mult) : raw[i];
                }
                if (dgi < 4u) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if ((unsigned)i == dgi) v[i] = fma(a.mult, __dadd_rn(raw[i], a.diag_rel), a.diag_abs);
                }
                if (full4) {
                    *reinterpret_cast<double2*>(colp) = make_double2(v[0], v[1]);
                    *reinterpret_cast<double2*>(colp + 2) = make_double2(v[2], v[3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (gi0 + i < a.nP) colp[i] = v[i];
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
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

bool assemble_fusable(int nterms, int nunits) { return nterms >= 1 && nterms <= FNT && nunits <= FCC; }

// symmetric training matrices of `ngroups` groups of batch elements; the elements of a group share structure and
// coordinates (groups == nullptr: one group of `count` elements).  The caller has checked assemble_fusable().
void launch_assemble_fused(const AsmArgs& a, const AsmGroup* groups, int ngroups, int count, cudaStream_t st) {
    if (a.nP <= 0 || ngroups <= 0) return;
    static std::atomic<int> nsm[FGP_MAXDEV];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= FGP_MAXDEV) dev = 0;
    if (!nsm[dev]) {
        int count = 0;
        cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev);
        nsm[dev] = count;
    }
    FusedK k{a, groups, count};
    const int nT = (a.nP + FR - 1) / FR;
    const int ntiles = (FR / FC) * nT * (nT + 1) / 2;
    // persistent CTAs, two per SM over the whole launch: every group gets its share of them (at least one, at most a CTA per tile)
    const int smCount = nsm[dev];
    const int slots = 2 * (smCount > 0 ? smCount : 148);
    const int gx = std::max(1, std::min(ntiles, (slots + ngroups - 1) / ngroups));
    dim3 grid(gx, ngroups);
    constexpr int dynBytes = 2 * FCC * (FR + FC) * (int)sizeof(double);
    static std::atomic<bool> attr_set[FGP_MAXDEV];        // zero-initialised; several host threads may arrive at once
    if (!attr_set[dev]) {
        cudaFuncSetAttribute(assemble_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, dynBytes);
        cudaFuncSetAttribute(assemble_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dynBytes);
        cudaFuncSetAttribute(assemble_fused_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, dynBytes);
        attr_set[dev] = true;
    }
    FGP_COUNT_LAUNCH();
    switch (a.ker) {
        case 1: assemble_fused_kernel<1><<<grid, FTHREADS, dynBytes, st>>>(k, ntiles); break;
        case 2: assemble_fused_kernel<2><<<grid, FTHREADS, dynBytes, st>>>(k, ntiles); break;
        default: assemble_fused_kernel<3><<<grid, FTHREADS, dynBytes, st>>>(k, ntiles); break;
    }
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
