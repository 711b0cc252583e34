"""Literal numpy/scipy restatement of funGp's GP hot path -- oracle/test infrastructure.

Every function cites the reference file:line (relative to /root/reference) it
follows.  The operation sequence is kept literal on purpose (materialised
per-input distance matrices, one temporary per R expression, product of
per-input exponentials, full Cholesky, dense inverse for LOOCV) because this
module doubles as the reported CPU baseline (`cpu_baseline.kind = "port"`).
All arrays are float64; matrices use numpy's (row, col) indexing of the same
mathematical objects R stores column-major.

NOT part of the product: see oracle/__init__.py.
"""
import math
import numpy as np
import scipy.linalg as sla
from scipy.interpolate import BSpline

KER_CODES = {"gauss": 1, "matern5_2": 2, "matern3_2": 3}      # R/1_Xfgpm_Class.R:457-459
DIST_CODES = {"L2_bygroup": 1, "L2_byindex": 2}              # R/3_ant_admin.R:301-373
BASIS_CODES = {"B-splines": 1, "PCA": 2}


# --------------------------------------------------------------------------------------
# dimension reduction -- R/7_dimRedFunctions.R
# --------------------------------------------------------------------------------------
def bspline_knots(k, p):
    """R/7_dimRedFunctions.R:40-49: order and clamped knot vector on the grid 1..k."""
    ord_ = p if p <= 3 else 4
    n_inner = p - ord_ + 2
    n_outer = ord_ - 1
    inner = np.linspace(1.0, float(k), n_inner)
    knots = np.concatenate([np.full(n_outer, 1.0), inner, np.full(n_outer, float(k))])
    return knots, ord_


def proj_bsplines(k, p):
    """R/7_dimRedFunctions.R:40-51: splineDesign(knots, x=1:k, outer.ok=TRUE, ord) -> k x p."""
    knots, ord_ = bspline_knots(k, p)
    x = np.arange(1, k + 1, dtype=np.float64)
    return np.asarray(BSpline.design_matrix(x, knots, ord_ - 1).todense())


def proj_pca(f, p):
    """R/7_dimRedFunctions.R:57-60: first p eigenvectors of cov(f), decreasing eigenvalue (sign = LAPACK's, Q7)."""
    C = np.cov(f, rowvar=False)
    w, V = np.linalg.eigh(C)
    return V[:, ::-1][:, :p].copy()


def project_coefs(B, f):
    """R/7_dimRedFunctions.R:18-19 (and R/1_fgpm_Class.R:844-845): t(solve(B'B, B' f'))."""
    Q = B.T @ B
    return np.linalg.solve(Q, B.T @ f.T).T, Q


def dimReduction(fIn, fpDims, methvec):
    """R/7_dimRedFunctions.R:7-29 -> dict(basis, coefs, J)."""
    basis, coefs, J = [], [], []
    for f, p, meth in zip(fIn, fpDims, methvec):
        f = np.asarray(f, dtype=np.float64)
        if p > 0:
            B = proj_bsplines(f.shape[1], p) if meth == "B-splines" else proj_pca(f, p)
            c, Q = project_coefs(B, f)
            basis.append(B); coefs.append(c); J.append(Q)
        else:
            I = np.eye(f.shape[1])
            basis.append(I); J.append(I.copy()); coefs.append(f.copy())
    return dict(basis=basis, coefs=coefs, J=J)


# --------------------------------------------------------------------------------------
# distances -- R/7_distanceFunctions.R
# --------------------------------------------------------------------------------------
def L2_byindex(s1, s2):
    """R/7_distanceFunctions.R:28-34: one |outer(-)| matrix per column."""
    return [np.abs(np.subtract.outer(s1[:, i], s2[:, i])) for i in range(s1.shape[1])]


def L2_bygroup(f1, f2, J):
    """R/7_distanceFunctions.R:38-46: sqrt(pmax(sum_a outer(f1_a,f2_a,-)*outer((f1 J)_a,(f2 J)_a,-), 0))."""
    f1J = f1 @ J
    f2J = f2 @ J
    Dm = np.zeros((f1.shape[0], f2.shape[0]))
    for a in range(J.shape[0]):
        Dm = Dm + np.subtract.outer(f1[:, a], f2[:, a]) * np.subtract.outer(f1J[:, a], f2J[:, a])
    return np.sqrt(np.maximum(Dm, 0.0))


def setDistMatrix_S(sIn1, sIn2):
    """R/7_distanceFunctions.R:20-22."""
    return L2_byindex(sIn1, sIn2)


def setDistMatrix_F(fIn1, fIn2, J, disType):
    """R/7_distanceFunctions.R:3-17."""
    Dl = []
    for i in range(len(J)):
        if disType[i] == "L2_byindex":
            Dl += L2_byindex(fIn1[i], fIn2[i])
        else:
            Dl.append(L2_bygroup(fIn1[i], fIn2[i], J[i]))
    return Dl


def getOwners(disType, p):
    """R/7_distanceFunctions.R:48-58 (1-based owner ids)."""
    owners = []
    for i, dt in enumerate(disType):
        owners += [i + 1] if dt == "L2_bygroup" else [i + 1] * int(p[i])
    return owners


# --------------------------------------------------------------------------------------
# correlation -- R/7_correlFunctions.R
# --------------------------------------------------------------------------------------
def gaussian_cor(Ms, thetas):
    """R/7_correlFunctions.R:31-37."""
    Rm = np.ones_like(Ms[0])
    for M, th in zip(Ms, thetas):
        Rm = Rm * np.exp(-0.5 * (M / th) ** 2)
    return Rm


def matern52_cor(Ms, thetas):
    """R/7_correlFunctions.R:53-60."""
    Rm = np.ones_like(Ms[0])
    s5 = math.sqrt(5.0)
    for M, th in zip(Ms, thetas):
        Rm = Rm * (1 + s5 * np.abs(M) / th + (5.0 / 3.0) * (M ** 2) / (th ** 2)) * np.exp(-s5 * np.abs(M / th))
    return Rm


def matern32_cor(Ms, thetas):
    """R/7_correlFunctions.R:74-80."""
    Rm = np.ones_like(Ms[0])
    s3 = math.sqrt(3.0)
    for M, th in zip(Ms, thetas):
        Rm = Rm * (1 + s3 * np.abs(M) / th) * np.exp(-s3 * np.abs(M / th))
    return Rm


def setR(thetas, Ms, kerType):
    """R/7_correlFunctions.R:1-15."""
    if len(Ms) == 0:
        raise ValueError("setR needs at least one distance matrix")
    return {"gauss": gaussian_cor, "matern5_2": matern52_cor, "matern3_2": matern32_cor}[kerType](Ms, thetas)


# --------------------------------------------------------------------------------------
# training -- R/3_training_{S,F,SF}.R
# --------------------------------------------------------------------------------------
class NotPosDef(Exception):
    """base::chol's 'the leading minor of order k is not positive definite'."""

    def __init__(self, order):
        super().__init__(f"the leading minor of order {order} is not positive definite")
        self.order = order


def r_chol_lower(A):
    """t(chol(A)) of base R (dpotrf); raises NotPosDef with the LAPACK info."""
    L, info = sla.lapack.dpotrf(A, lower=1, clean=1)
    if info > 0:
        raise NotPosDef(info)
    return L


def corr_matrix(thetas, Ms, kerType, nugget):
    """R = setR(thetas_s, sMs) * setR(thetas_f, fMs) + diag(nugget) (R/3_training_SF.R:248).

    The scalar/functional split is a product of two products, so one product over the
    concatenated list is the same arithmetic up to the order of multiplications.
    """
    R = setR(thetas, Ms, kerType)
    n = R.shape[0]
    return R + np.diag(np.full(n, nugget))


def analyticVar_llik(L, y, n):
    """R/3_training_SF.R:266-269: crossprod(backsolve(t(U), y, upper.tri=FALSE)) / n."""
    z = sla.solve_triangular(L, y, lower=True)
    return float(z @ z) / n


def negLogLik(thetas, sMs, fMs, y, kerType, nugget, var_known=None):
    """R/3_training_SF.R:231-258 (also _S.R:200-213, _F.R:197-210); Q1: '+ n' even with known variance."""
    y = np.asarray(y, dtype=np.float64).ravel()
    n = y.size
    ds = len(sMs)
    if ds and len(fMs):
        R = setR(thetas[:ds], sMs, kerType) * setR(thetas[ds:], fMs, kerType)
    elif ds:
        R = setR(thetas, sMs, kerType)
    else:
        R = setR(thetas, fMs, kerType)
    R = R + np.diag(np.full(n, nugget))
    L = r_chol_lower(R)
    sig2 = analyticVar_llik(L, y, n) if var_known is None else float(var_known)
    llik = -0.5 * (n * math.log(2 * math.pi * sig2) + 2 * np.sum(np.log(np.diag(L))) + n)
    return -llik, sig2


def setBounds(Ms):
    """R/3_training_SF.R:57-74: lower 1e-10, upper 2*max(D_l)."""
    ul = np.array([2.0 * M.max() for M in Ms])
    return np.vstack([np.full(len(Ms), 1e-10), ul])


def r_order(x):
    """R's order(): stable, ascending; NaN last."""
    return np.argsort(np.asarray(x), kind="stable")


def setSPoints(bnds, nll_fn, n_starts, n_presample, rng, spoints_usr=None):
    """R/3_training_SF.R:83-108.  rng.runif consumes n.ls*n.presample draws, column-major fill."""
    ll, ul = bnds[0], bnds[1]
    n_ls = bnds.shape[1]
    u = rng.runif(n_ls * n_presample).reshape((n_presample, n_ls)).T  # matrix(nrow=n.ls, ncol=n.presample)
    allsp = ll[:, None] + u * (ul - ll)[:, None]
    fit = np.array([nll_fn(allsp[:, j]) for j in range(n_presample)])
    sp = allsp[:, r_order(fit)[:n_starts]].copy()
    if spoints_usr is not None:
        n_usr = spoints_usr.shape[1]
        n_rep = min(n_starts, n_usr)
        sp[:, n_starts - n_rep:n_starts] = spoints_usr
    return sp, allsp, fit


def r_fd_gradient(fn, x, lower, upper, ndeps=1e-3):
    """R's optim() numerical gradient (src/library/stats/src/optim.c fmingr, not vendored):
    central differences with step ndeps, clipped at the box."""
    g = np.empty_like(x)
    for i in range(x.size):
        eps = ndeps
        xp = x.copy(); xp[i] = min(x[i] + eps, upper[i]); epsused = xp[i] - x[i]
        v1 = fn(xp)
        xm = x.copy(); xm[i] = max(x[i] - eps, lower[i]); eps2 = x[i] - xm[i]
        v2 = fn(xm)
        g[i] = (v1 - v2) / (epsused + eps2)
    return g


def optim_lbfgsb(fn, x0, lower, upper, maxit=100, factr=1e7, pgtol=0.0, lmm=5):
    """Stand-in for stats::optim(method='L-BFGS-B') (R/3_training_SF.R:125): scipy's L-BFGS-B with
    R's control defaults and R's finite-difference gradient.  Not R's own v2.x port."""
    from scipy.optimize import minimize
    res = minimize(fn, x0, jac=lambda x: r_fd_gradient(fn, x, lower, upper), method="L-BFGS-B",
                   bounds=list(zip(lower, upper)),
                   options=dict(maxcor=lmm, ftol=factr * np.finfo(float).eps, gtol=pgtol, maxiter=maxit,
                                maxfun=15000, maxls=20))
    return res.x, float(res.fun), (0 if res.status == 0 else (1 if res.status == 1 else 52))


# --------------------------------------------------------------------------------------
# prediction / simulation -- R/4_prediction_*.R, R/5_simulation_*.R
# --------------------------------------------------------------------------------------
def preMats(Ms, y, sig2, thetas, kerType, nugget):
    """R/4_prediction_SF.R:24-32."""
    R = setR(thetas, Ms, kerType)
    n = R.shape[0]
    K = sig2 * (R + np.diag(np.full(n, nugget)))
    L = r_chol_lower(K)
    LInvY = sla.solve_triangular(L, np.asarray(y, dtype=np.float64).ravel(), lower=True)
    return L, LInvY


def makePreds(Ms_tp, Ms_pp, sig2, thetas, kerType, L, LInvY, detail, nugget):
    """R/4_prediction_SF.R:2-21."""
    K_tp = sig2 * setR(thetas, Ms_tp, kerType)
    LInvK = sla.solve_triangular(L, K_tp, lower=True)
    out = dict(mean=LInvK.T @ LInvY,
               sd=np.sqrt(np.maximum(sig2 - np.sum(LInvK * LInvK, axis=0), 0.0)))
    if detail == "full":
        out["K.tp"] = K_tp
        R = setR(thetas, Ms_pp, kerType)
        out["K.pp"] = sig2 * (R + np.diag(np.full(R.shape[0], nugget)))
    return out


def makeSims(Ms_ts, Ms_ss, sig2, thetas, kerType, L, LInvY, Z, nugget_sm, detail):
    """R/5_simulation_SF.R:3-23; Z = matrix(rnorm(n.sm*nsim), n.sm, nsim) is passed in."""
    K_ts = sig2 * setR(thetas, Ms_ts, kerType)
    K_ss = sig2 * setR(thetas, Ms_ss, kerType)
    LInvK = sla.solve_triangular(L, K_ts, lower=True)
    mean = LInvK.T @ LInvY
    n_sm = K_ss.shape[0]
    C = K_ss - LInvK.T @ LInvK + np.diag(np.full(n_sm, nugget_sm))
    Lc = r_chol_lower(C)
    noise = Lc @ Z
    out = dict(sims=(mean[:, None] + noise).T)
    if detail == "full":
        out["mean"] = mean
        out["sd"] = np.sqrt(np.maximum(sig2 - np.sum(LInvK * LInvK, axis=0), 0.0))
    return out


# --------------------------------------------------------------------------------------
# LOOCV fitness -- R/8_outilsStats.R:1-7, R/1_Xfgpm_Class.R:522-552
# --------------------------------------------------------------------------------------
def getOut_loocv(L, sig2, nugget, y):
    """R/8_outilsStats.R:1-7 (Q2: the nugget ends up counted twice)."""
    n = L.shape[0]
    R = (L @ L.T) / sig2 + np.diag(np.full(n, nugget))
    Rinv = np.linalg.inv(R)
    y = np.asarray(y, dtype=np.float64).ravel()
    return y - (Rinv @ y) / np.diag(Rinv)


def q2(y, yhat):
    """R/1_Xfgpm_Class.R:541, 548."""
    y = np.asarray(y, dtype=np.float64).ravel()
    return 1.0 - np.mean((y - yhat) ** 2) / np.mean((y - np.mean(y)) ** 2)


# --------------------------------------------------------------------------------------
# black boxes -- R/7_blackBoxFunctions.R
# --------------------------------------------------------------------------------------
def fgp_BB3(sIn, fIn):
    """R/7_blackBoxFunctions.R:139-150."""
    f1, f2 = fIn
    t1 = np.linspace(0, 1, f1.shape[1])
    return sIn[:, 0] + 2 * sIn[:, 1] + 4 * np.mean(t1 * f1, axis=1) + np.mean(f2, axis=1)


def fgp_BB7(sIn, fIn):
    """R/7_blackBoxFunctions.R:219-239."""
    x1, x2, x3, x4, x5 = (sIn[:, i] for i in range(5))
    f1, f2 = fIn
    t1 = np.linspace(0, 1, f1.shape[1]); t2 = np.linspace(0, 1, f2.shape[1])
    a = (x2 + 4 * x3 - (5 / (4 * math.pi ** 2)) * x1 ** 2 + (5 / math.pi) * x1 - 6) ** 2
    b = 10 * (1 - (1 / (8 * math.pi))) * np.cos(x1) * x2 ** 2 * x5 ** 3
    c = 10
    d1 = 42 * np.sin(x4) * np.mean(15 * f1 * (1 - t1) - 5, axis=1)
    d2 = math.pi * (((x1 * x5 + 5) / 5) + 15) * np.mean(15 * t2 * f2, axis=1)
    d = (4 / 3) * math.pi * (d1 + d2)
    return a + b + c + d


# --------------------------------------------------------------------------------------
# whole-model restatement -- R/1_fgpm_Class.R:345-585 (numerics only)
# --------------------------------------------------------------------------------------
class RefModel:
    """Numeric content of an S4 `fgpm` object (R/1_fgpm_Class.R:65-84)."""


def fgpm(sIn=None, fIn=None, sOut=None, kerType="matern5_2", f_disType="L2_bygroup", f_pdims=3,
         f_basType="B-splines", var_hyp=None, ls_s_hyp=None, ls_f_hyp=None, nugget=1e-8,
         n_starts=1, n_presample=20, rng=None, spoints_usr=None):
    """R/1_fgpm_Class.R:345-585: projection -> distances -> setHypers_* -> preMats_*.

    `rng` is an oracle.rrng.RRng positioned where R's stream would be when fgpm() is called.
    Partial-known length-scales (cases 2/3 of R/3_training_SF.R:235-244) are supported.
    """
    m = RefModel()
    y = np.asarray(sOut, dtype=np.float64).ravel()
    n = y.size
    n_presample = max(n_presample, n_starts)                     # R/1_fgpm_Class.R:372
    m.sIn = None if sIn is None else np.asarray(sIn, dtype=np.float64).reshape(n, -1)
    m.ds = 0 if m.sIn is None else m.sIn.shape[1]
    m.fIn = None
    m.df = 0
    sMs, fMs = [], []
    if fIn is not None:
        if isinstance(fIn, np.ndarray):
            fIn = [fIn]
        m.fIn = [np.asarray(f, dtype=np.float64) for f in fIn]
        m.df = len(m.fIn)
        rep = lambda v: list(v) if isinstance(v, (list, tuple, np.ndarray)) else [v] * m.df
        m.f_disType, m.f_pdims, m.f_basType = rep(f_disType), [int(p) for p in rep(f_pdims)], rep(f_basType)
        bcj = dimReduction(m.fIn, m.f_pdims, m.f_basType)
        m.basis, m.coefs, m.J = bcj["basis"], bcj["coefs"], bcj["J"]
        fMs = setDistMatrix_F(m.coefs, m.coefs, m.J, m.f_disType)
        m.owners = getOwners(m.f_disType, [c.shape[1] for c in m.coefs])
    if m.sIn is not None:
        sMs = setDistMatrix_S(m.sIn, m.sIn)
    m.kerType, m.nugget, m.y, m.n = kerType, nugget, y, n
    d_s, d_f = len(sMs), len(fMs)
    ls_s_known = None if ls_s_hyp is None else np.asarray(ls_s_hyp, dtype=np.float64)
    ls_f_known = None if ls_f_hyp is None else np.asarray(ls_f_hyp, dtype=np.float64)
    s_free = d_s > 0 and ls_s_known is None
    f_free = d_f > 0 and ls_f_known is None

    def full_theta(th):
        parts = []
        if d_s:
            parts.append(th[:d_s] if s_free and f_free else (th if s_free else ls_s_known))
        if d_f:
            parts.append(th[d_s:] if s_free and f_free else (th if f_free else ls_f_known))
        return np.concatenate(parts)

    Ms = sMs + fMs
    m.convergence = float("nan"); m.negLogLik = float("nan")
    if not s_free and not f_free:
        theta = full_theta(np.zeros(0))
        if var_hyp is None:                                       # R/3_training_SF.R:7-18
            R = corr_matrix(theta, Ms, kerType, nugget)
            sig2 = analyticVar_llik(r_chol_lower(R), y, n)
        else:
            sig2 = float(var_hyp)
    else:
        free_Ms = (sMs if s_free else []) + (fMs if f_free else [])
        bnds = setBounds(free_Ms)
        nll_fn = lambda th: negLogLik(full_theta(np.asarray(th)), [], Ms, y, kerType, nugget, var_hyp)[0]
        sp, _, _ = setSPoints(bnds, nll_fn, n_starts, n_presample, rng, spoints_usr)
        best = None
        for i in range(n_starts):                                 # R/3_training_SF.R:140-158
            try:
                x, f, conv = optim_lbfgsb(nll_fn, sp[:, i], bnds[0], bnds[1])
            except NotPosDef:
                continue
            if best is None or f < best[1]:
                best = (x, f, conv)
        if best is None:
            raise RuntimeError("All model optimizations crashed!")
        theta = full_theta(best[0])
        m.negLogLik, m.convergence = best[1], best[2]
        R = corr_matrix(theta, Ms, kerType, nugget)               # R/3_training_SF.R:215-220
        sig2 = analyticVar_llik(r_chol_lower(R), y, n) if var_hyp is None else float(var_hyp)
    m.varHyp, m.lsHyps = sig2, theta
    m.s_lsHyps, m.f_lsHyps = theta[:d_s], theta[d_s:]
    m.L, m.LInvY = preMats(Ms, y, sig2, theta, kerType, nugget)
    return m


def _new_dists(m, sNew, fNew):
    Ms_tp, Ms_pp = [], []
    if m.ds:
        sNew = np.asarray(sNew, dtype=np.float64).reshape(-1, m.ds)
        Ms_tp += setDistMatrix_S(m.sIn, sNew); Ms_pp += setDistMatrix_S(sNew, sNew)
    if m.df:
        if isinstance(fNew, np.ndarray):
            fNew = [fNew]
        cN = [project_coefs(B, np.asarray(f, dtype=np.float64))[0] for B, f in zip(m.basis, fNew)]  # R/1_fgpm_Class.R:844-845
        J = [B.T @ B for B in m.basis]
        Ms_tp += setDistMatrix_F(m.coefs, cN, J, m.f_disType); Ms_pp += setDistMatrix_F(cN, cN, J, m.f_disType)
    return Ms_tp, Ms_pp


QNORM975 = 1.959963984540054  # qnorm(0.975), R/1_fgpm_Class.R:888


def predict(m, sNew=None, fNew=None, detail="light"):
    """R/1_fgpm_Class.R:825-909."""
    Ms_tp, Ms_pp = _new_dists(m, sNew, fNew)
    out = makePreds(Ms_tp, Ms_pp, m.varHyp, m.lsHyps, m.kerType, m.L, m.LInvY, detail, m.nugget)
    out["lower95"] = out["mean"] - QNORM975 * out["sd"]
    out["upper95"] = out["mean"] + QNORM975 * out["sd"]
    return out


def simulate(m, Z, sNew=None, fNew=None, nugget_sm=0.0, detail="light"):
    """R/1_fgpm_Class.R:1026-1116; Z (n.sm x nsim) = matrix(rnorm(n.sm*nsim), n.sm, nsim) drawn by the caller."""
    Ms_ts, Ms_ss = _new_dists(m, sNew, fNew)
    out = makeSims(Ms_ts, Ms_ss, m.varHyp, m.lsHyps, m.kerType, m.L, m.LInvY, Z, nugget_sm, detail)
    if detail == "full":
        out["lower95"] = out["mean"] - QNORM975 * out["sd"]
        out["upper95"] = out["mean"] + QNORM975 * out["sd"]
    return out


def q2_loocv(m):
    """getFitness(model) with stat Q2loocv, R/1_Xfgpm_Class.R:522-543."""
    return q2(m.y, getOut_loocv(m.L, m.varHyp, m.nugget, m.y))
