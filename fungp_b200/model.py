"""Host-side mirror of the reference's model API for the hot path: fgpm(), predict(), simulate(), update().

Same argument names, meaning and error behaviour as the R functions (R/1_fgpm_Class.R:345-585, 819-909, 1018-1116,
1320-1506; R/6_updating.R) -- in the R package these stay R and call the C ABI through rglue/; this module is the
equivalent caller for an environment without R.  Every number is produced by libfungp_b200.so on the GPU; the only
host arithmetic here is what stays in R in the real integration (argument handling, the L-BFGS-B driver, the 95 % bands).

The optimiser: the reference calls stats::optim(method = "L-BFGS-B") without a gradient, so R differentiates by central
finite differences (ndeps = 1e-3, clipped at the box).  Here one optimiser iteration = ONE batched device call that
evaluates theta and its 2d finite-difference neighbours (fgp_nll_batch), which is what the R glue passes as `gr=`.
"""
import math

import numpy as np

from .core import Context, Problem, NotPosDefError, FgpError, KER, DIST, BASIS  # noqa: F401

QNORM975 = 1.959963984540054   # qnorm(0.975), R/1_fgpm_Class.R:888


class NumpyRng:
    """Default stand-in for R's RNG stream (runif / rnorm); tests pass an R-compatible generator instead."""

    def __init__(self, seed=None):
        self.g = np.random.default_rng(seed)

    def runif(self, n):
        return self.g.uniform(size=int(n))

    def rnorm(self, n):
        return self.g.standard_normal(int(n))


class FgpProj:
    """fgpProj (R/2_fgpProj_Class.R:19-28)"""

    def __init__(self, pdims, basType, basis, coefs, J):
        self.pdims, self.basType, self.basis, self.coefs, self.J = pdims, basType, basis, coefs, J


class FgpKern:
    """fgpKern (R/2_fgpKern_Class.R:23-32)"""

    def __init__(self):
        self.kerType = None; self.f_disType = []; self.varHyp = None
        self.s_lsHyps = np.zeros(0); self.f_lsHyps = np.zeros(0); self.f_lsOwners = []


class Fgpm:
    """Numeric content of the S4 class `fgpm` (R/1_fgpm_Class.R:65-84) + the device problem handle."""

    def close(self):
        if getattr(self, "_problem", None) is not None:
            self._problem.close()
            self._problem = None


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None or _default_ctx.h is None:
        _default_ctx = Context(0)     # all visible GPUs; batched calls shard over them
    return _default_ctx


def _rep(v, df):
    return list(v) if isinstance(v, (list, tuple, np.ndarray)) else [v] * df


def r_fd_points(x, lower, upper, ndeps=1e-3):
    """columns: x, then for every i: x + eps e_i (clipped), x - eps e_i (clipped)  -- optim's fmingr"""
    d = x.size
    pts = np.repeat(x[:, None], 2 * d + 1, axis=1)
    for i in range(d):
        pts[i, 1 + 2 * i] = min(x[i] + ndeps, upper[i])
        pts[i, 2 + 2 * i] = max(x[i] - ndeps, lower[i])
    return pts


def optim_lbfgsb_batched(nll_batch, x0, lower, upper, maxit=100, factr=1e7, pgtol=0.0, lmm=5):
    """L-BFGS-B with R's control defaults (R/3_training_SF.R:125; optim: lmm 5, factr 1e7, pgtol 0, maxit 100) and R's
    finite-difference gradient, value and gradient coming from one batched device call per iterate.
    nll_batch(points d x B) -> nll[B]; it raises NotPosDefError when the iterate or any of its finite-difference neighbours
    cannot be evaluated (R would stop inside fn / gr), which the caller treats as a lost start."""
    from scipy.optimize import minimize
    lower = np.asarray(lower, dtype=np.float64); upper = np.asarray(upper, dtype=np.float64)

    def fg(x):
        pts = r_fd_points(np.asarray(x, dtype=np.float64), lower, upper)
        v = nll_batch(pts)
        d = x.size
        g = np.empty(d)
        for i in range(d):
            g[i] = (v[1 + 2 * i] - v[2 + 2 * i]) / (pts[i, 1 + 2 * i] - pts[i, 2 + 2 * i])
        return float(v[0]), g

    res = minimize(fg, np.asarray(x0, dtype=np.float64), jac=True, method="L-BFGS-B", bounds=list(zip(lower, upper)),
                   options=dict(maxcor=lmm, ftol=factr * np.finfo(float).eps, gtol=pgtol, maxiter=maxit, maxfun=15000,
                                maxls=20))
    return res.x, float(res.fun), (0 if res.status == 0 else (1 if res.status == 1 else 52))


def fgpm(sIn=None, fIn=None, sOut=None, kerType="matern5_2", f_disType="L2_bygroup", f_pdims=3, f_basType="B-splines",
         var_hyp=None, ls_s_hyp=None, ls_f_hyp=None, nugget=1e-8, n_starts=1, n_presample=20, rng=None, spoints_usr=None,
         ctx=None, want_L=True, reuse=None, optimizer="host"):
    """fgpm() (R/1_fgpm_Class.R:345-585): projection -> device problem -> setHypers_* -> preMats_*.
    optimizer: "host" drives scipy's L-BFGS-B from this process, one batched device call per iterate (the role stats::optim
    keeps in the R package); "lockstep" hands the whole multistart fit to the library's own L-BFGS-B (fgp_fit_batch with this
    structure as its only candidate): all n_starts starts advance together, one launch sequence per round -- available when
    no hyper-parameter is given and no user start points are.  Same presample draws from `rng` either way.
    reuse: a fitted model whose factor may be re-used (update() fast path): honoured when every hyper-parameter is given
    and the length-scales, kernel and nugget are those of `reuse` -- the leading points both data sets share keep their
    block of the factor (fgp_premats_reuse), a different variance is applied by rescaling (fgp_update_var)."""
    if sOut is None or (sIn is None and fIn is None):
        raise ValueError("fgpm: sOut and at least one of sIn / fIn are required")      # checkVal_fgpm
    if kerType not in KER:
        raise ValueError(f"fgpm: unknown kerType {kerType!r}")
    ctx = ctx or default_context()
    rng = rng or NumpyRng()
    m = Fgpm()
    y = np.asarray(sOut, dtype=np.float64).ravel()
    n = y.size
    n_presample = max(n_presample, n_starts)                        # R/1_fgpm_Class.R:372
    m.sOut, m.n_tot, m.n_tr, m.nugget = y, n, n, nugget
    m.sIn = None if sIn is None else np.asarray(sIn, dtype=np.float64).reshape(n, -1)
    m.ds = 0 if m.sIn is None else m.sIn.shape[1]
    m.kern = FgpKern(); m.kern.kerType = kerType
    m.fIn, m.df, m.f_proj = None, 0, None
    coefs, J, dis = [], [], []
    if fIn is not None:
        if isinstance(fIn, np.ndarray):
            fIn = [fIn]
        m.fIn = [np.asarray(f, dtype=np.float64) for f in fIn]
        m.df = len(m.fIn)
        dis, pdims, bas = _rep(f_disType, m.df), [int(p) for p in _rep(f_pdims, m.df)], _rep(f_basType, m.df)
        basis = []
        for f, p, b in zip(m.fIn, pdims, bas):                     # dimReduction, R/7_dimRedFunctions.R:7-29
            Bm, Jm, co = ctx.project(f, p, b if b in BASIS else "B-splines")
            basis.append(Bm); J.append(Jm); coefs.append(co)
        m.f_proj = FgpProj(pdims, bas, basis, coefs, J)
        m.kern.f_disType = dis
        owners = []
        for i, (dt, c) in enumerate(zip(dis, coefs)):              # getOwners, R/7_distanceFunctions.R:48-58
            owners += [i + 1] if dt == "L2_bygroup" else [i + 1] * c.shape[1]
        m.kern.f_lsOwners = owners
    m.type = "hybrid" if (m.ds and m.df) else ("functional" if m.df else "scalar")
    P = Problem(ctx, m.sIn, coefs, J, dis, y)
    m._problem, m._ctx = P, ctx
    d_s, d_f = m.ds, P.d - m.ds
    ls_s_known = None if ls_s_hyp is None else np.asarray(ls_s_hyp, dtype=np.float64).ravel()
    ls_f_known = None if ls_f_hyp is None else np.asarray(ls_f_hyp, dtype=np.float64).ravel()
    s_free = d_s > 0 and ls_s_known is None
    f_free = d_f > 0 and ls_f_known is None

    def full_theta(th):                                            # the three cases of R/3_training_SF.R:235-244
        th = np.asarray(th, dtype=np.float64)
        if th.ndim == 1:
            th = th[:, None]
        parts = []
        if d_s:
            parts.append(th[:d_s] if (s_free and f_free) else (th if s_free else np.repeat(ls_s_known[:, None], th.shape[1], 1)))
        if d_f:
            parts.append(th[d_s:] if (s_free and f_free) else (th if f_free else np.repeat(ls_f_known[:, None], th.shape[1], 1)))
        return np.vstack(parts)

    m.convergence, m.negLogLik = float("nan"), float("nan")
    if not s_free and not f_free:                                  # R/3_training_SF.R:7-18
        theta = full_theta(np.zeros((0, 1)))[:, 0]
        if var_hyp is None:
            nll, s2, info = P.nll_batch(kerType, nugget, theta)
            if info[0]:
                raise NotPosDefError(int(info[0]))
            sig2 = float(s2[0])
        else:
            sig2 = float(var_hyp)
    else:
        dmax = P.dist_max()                                        # setBounds_*, R/3_training_SF.R:57-74
        free = np.r_[np.arange(d_s) if s_free else [], (d_s + np.arange(d_f)) if f_free else []].astype(int)
        lower, upper = np.full(free.size, 1e-10), 2.0 * dmax[free]

        def nll_pts(pts, strict):
            v, _, info = P.nll_batch(kerType, nugget, full_theta(pts), var_known=var_hyp)
            if strict and info.any():                              # quirk Q6: presampling does not catch chol errors
                raise NotPosDefError(int(info[np.nonzero(info)[0][0]]))
            return v

        # setSPoints_*, R/3_training_SF.R:83-108: ONE batched call for the n.presample points
        u_flat = np.asarray(rng.runif(free.size * n_presample))
        if optimizer == "lockstep" and s_free == (d_s > 0) and f_free == (d_f > 0) and var_hyp is None and spoints_usr is None:
            from . import factory as F
            ant = F.encode_ant(m.ds, m.df, list(range(m.ds)), list(range(m.df)), dis, pdims if m.df else [],
                               [b if b in BASIS else "B-splines" for b in bas] if m.df else [], kerType)
            r = ctx.fit_batch(m.sIn, m.fIn, y, ant[None, :], [u_flat], nugget=nugget, n_starts=n_starts, n_presample=n_presample,
                              dmax=P.d)
            if r["info"][0] > 0:
                raise NotPosDefError(int(r["info"][0]))
            if r["info"][0] < 0:
                raise RuntimeError("All model optimizations crashed!")
            theta = r["theta"][0][:P.d].copy()
            m.negLogLik, m.convergence = float(r["nll"][0]), 0.0
            m.kern.varHyp = float(r["sig2"][0])
            m.kern.s_lsHyps, m.kern.f_lsHyps = theta[:d_s].copy(), theta[d_s:].copy()
            m.reused_points = 0
            L, z = P.premats(kerType, nugget, m.kern.varHyp, theta, want_L=want_L)
            m.preMats = dict(L=L, LInvY=z)
            return m
        u = u_flat.reshape((n_presample, free.size)).T
        allsp = lower[:, None] + u * (upper - lower)[:, None]
        fit = nll_pts(allsp, strict=True)
        sp = allsp[:, np.argsort(fit, kind="stable")[:n_starts]].copy()
        if spoints_usr is not None:
            su = np.asarray(spoints_usr, dtype=np.float64).reshape(free.size, -1)
            n_rep = min(n_starts, su.shape[1])
            sp[:, n_starts - n_rep:n_starts] = su[:, :n_rep]
        best = None
        for i in range(n_starts):                                  # optimHypers_*, R/3_training_SF.R:122-223
            try:
                def nb(pts):
                    # any point of the finite-difference batch that cannot be evaluated is an error inside optim(): the start
                    # is lost (tryCatch below), exactly as in fitbatch.cu / fitlock.cu and in R
                    v, _, info = P.nll_batch(kerType, nugget, full_theta(pts), var_known=var_hyp)
                    if info.any():
                        raise NotPosDefError(int(info[np.nonzero(info)[0][0]]))
                    if not np.all(np.isfinite(v)):
                        raise ValueError("non-finite likelihood")
                    return v
                x, f, conv = optim_lbfgsb_batched(nb, sp[:, i], lower, upper)
            except (NotPosDefError, ValueError):
                continue                                           # tryCatch per start, :141-153
            if best is None or f < best[1]:
                best = (x, f, conv)
        if best is None:
            raise RuntimeError("All model optimizations crashed!")  # :189
        theta = full_theta(best[0])[:, 0]
        m.negLogLik, m.convergence = best[1], best[2]
        if var_hyp is None:                                        # recompute at the optimum, :215-220
            _, s2, info = P.nll_batch(kerType, nugget, theta)
            if info[0]:
                raise NotPosDefError(int(info[0]))
            sig2 = float(s2[0])
        else:
            sig2 = float(var_hyp)
    m.kern.varHyp = sig2
    m.kern.s_lsHyps, m.kern.f_lsHyps = theta[:d_s].copy(), theta[d_s:].copy()
    m.reused_points = 0
    ro = None if reuse is None else getattr(reuse, "_problem", None)
    if (ro is not None and ro.h and not s_free and not f_free and var_hyp is not None and reuse.kern.kerType == kerType
            and reuse.nugget == nugget and ro.d == P.d and ro.ctx is ctx
            and np.array_equal(np.r_[reuse.kern.s_lsHyps, reuse.kern.f_lsHyps], theta)):
        try:
            L, z, m.reused_points = P.premats_reuse(ro, want_L=want_L and sig2 == reuse.kern.varHyp)
            if sig2 != reuse.kern.varHyp:
                L, z = P.update_var(sig2, want_L=want_L)
            m.preMats = dict(L=L, LInvY=z)
            return m
        except FgpError as e:
            if isinstance(e, NotPosDefError):
                raise
            # structures differ (another projection dimension, distance or Gram matrix): plain preMats below
    L, z = P.premats(kerType, nugget, sig2, theta, want_L=want_L)  # preMats_*, R/4_prediction_SF.R:24-32
    m.preMats = dict(L=L, LInvY=z)
    return m


def _project_new(model, fNew):
    if isinstance(fNew, np.ndarray):
        fNew = [fNew]
    out = []
    for f, p, Bm in zip(fNew, model.f_proj.pdims, model.f_proj.basis):     # R/1_fgpm_Class.R:844-845
        out.append(model._ctx.project(np.asarray(f, dtype=np.float64), p, B=Bm)[2])
    return out


def _check_new(model, sIn, fIn, what):
    if model.ds and sIn is None:
        raise ValueError(f"{what}: scalar inputs are required for this model")         # checkVal_pred_and_sim
    if model.df and fIn is None:
        raise ValueError(f"{what}: functional inputs are required for this model")
    m = (np.asarray(sIn).reshape(-1, model.ds).shape[0] if model.ds else
         (fIn[0] if isinstance(fIn, (list, tuple)) else fIn).shape[0])
    return m


def predict(model, sIn_pr=None, fIn_pr=None, detail="light"):
    """predict.fgpm (R/1_fgpm_Class.R:825-909) -> dict(mean, sd, lower95, upper95 [, K.tp, K.pp])"""
    m = _check_new(model, sIn_pr, fIn_pr, "predict")
    cn = _project_new(model, fIn_pr) if model.df else []
    out = model._problem.predict(m, sIn_pr if model.ds else None, cn, detail=detail)
    out["lower95"] = out["mean"] - QNORM975 * out["sd"]                       # :888-889
    out["upper95"] = out["mean"] + QNORM975 * out["sd"]
    return out


def simulate(model, nsim=1, seed=None, sIn_sm=None, fIn_sm=None, nugget_sm=0.0, detail="light", rng=None, Z=None):
    """simulate.fgpm (R/1_fgpm_Class.R:1026-1116).  The normal draws are made on the host by the caller's generator
    (`rng.rnorm`, seeded as R does with set.seed(seed) right before the draw, R/5_simulation_SF.R:13-14) or passed as Z."""
    m = _check_new(model, sIn_sm, fIn_sm, "simulate")
    cn = _project_new(model, fIn_sm) if model.df else []
    if Z is None:
        rng = rng or NumpyRng(seed)
        Z = np.asarray(rng.rnorm(m * nsim)).reshape((nsim, m)).T            # matrix(rnorm(n.sm*nsim), n.sm, nsim)
    res = model._problem.simulate(m, Z, sIn_sm if model.ds else None, cn, nugget_sm=nugget_sm)
    if detail == "light":
        return res["sims"]
    res["lower95"] = res["mean"] - QNORM975 * res["sd"]                        # :1094-1095
    res["upper95"] = res["mean"] + QNORM975 * res["sd"]
    return res


def loocv_q2(model):
    """getFitness(model) with Q2loocv (R/1_Xfgpm_Class.R:522-543; getOut_loocv R/8_outilsStats.R:1-7)."""
    return model._problem.loocv()[1]


def update(model, sIn_nw=None, fIn_nw=None, sOut_nw=None, sIn_sb=None, fIn_sb=None, sOut_sb=None, ind_sb=None, ind_dl=None,
           var_sb=None, ls_s_sb=None, ls_f_sb=None, var_re=False, ls_s_re=False, ls_f_re=False, rng=None, ctx=None):
    """update.fgpm (R/1_fgpm_Class.R:1335-1506, R/6_updating.R): data deletion / substitution / addition, hyper-parameter
    substitution or re-estimation.  As in the reference every branch ends in an fgpm() call with the retained
    hyper-parameters passed as known (R/6_updating.R:7-89, 91-323, 325-500, 502-565, 567-643); indices are 1-based.
    Where the length-scales are retained that call re-uses the model's factor (fgpm(reuse=model)): the points in front of
    the first changed one keep their block, a new variance is a rescale -- instead of the reference's full refit."""
    sIn = None if model.sIn is None else model.sIn.copy()
    fIn = None if model.fIn is None else [f.copy() for f in model.fIn]
    y = model.sOut.copy()
    if ind_dl is not None:                                         # upd_del
        keep = np.setdiff1d(np.arange(y.size), np.asarray(ind_dl, dtype=int) - 1)
        sIn = None if sIn is None else sIn[keep]
        fIn = None if fIn is None else [f[keep] for f in fIn]
        y = y[keep]
    if ind_sb is not None:                                         # upd_subData
        idx = np.asarray(ind_sb, dtype=int) - 1
        if sIn_sb is not None:
            sIn[idx] = np.asarray(sIn_sb, dtype=np.float64).reshape(idx.size, -1)
        if fIn_sb is not None:
            for f, fs in zip(fIn, fIn_sb if isinstance(fIn_sb, (list, tuple)) else [fIn_sb]):
                f[idx] = fs
        if sOut_sb is not None:
            y[idx] = np.asarray(sOut_sb, dtype=np.float64).ravel()
    if sOut_nw is not None:                                        # upd_add
        if sIn is not None:
            sIn = np.vstack([sIn, np.asarray(sIn_nw, dtype=np.float64).reshape(-1, sIn.shape[1])])
        if fIn is not None:
            fIn = [np.vstack([f, fn]) for f, fn in zip(fIn, fIn_nw if isinstance(fIn_nw, (list, tuple)) else [fIn_nw])]
        y = np.r_[y, np.asarray(sOut_nw, dtype=np.float64).ravel()]
    k = model.kern
    var = None if var_re else (k.varHyp if var_sb is None else float(var_sb))
    ls_s = None if (ls_s_re or not model.ds) else (k.s_lsHyps if ls_s_sb is None else np.asarray(ls_s_sb, dtype=np.float64))
    ls_f = None if (ls_f_re or not model.df) else (k.f_lsHyps if ls_f_sb is None else np.asarray(ls_f_sb, dtype=np.float64))
    if var is not None and (ls_s_re or ls_f_re) and not var_re and var_sb is None:
        var = None                                                 # re-estimating length-scales re-estimates the variance too
    proj = model.f_proj
    return fgpm(sIn=sIn, fIn=fIn, sOut=y, kerType=k.kerType, f_disType=k.f_disType or "L2_bygroup",
                f_pdims=proj.pdims if proj else 3, f_basType=proj.basType if proj else "B-splines", var_hyp=var, ls_s_hyp=ls_s,
                ls_f_hyp=ls_f, nugget=model.nugget, rng=rng, ctx=ctx or model._ctx, reuse=model)
