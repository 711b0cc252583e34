"""Thin numpy-facing wrappers over the C ABI: Context and Problem.

Everything numeric happens in libfungp_b200.so on the GPU; these classes only marshal column-major
float64 arrays (R's layout) in and out.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib

KER = {"gauss": 1, "matern5_2": 2, "matern3_2": 3}
DIST = {"L2_bygroup": 1, "L2_byindex": 2}
BASIS = {"B-splines": 1, "PCA": 2}


class FgpError(RuntimeError):
    pass


class NotPosDefError(FgpError):
    """base::chol's error: 'the leading minor of order k is not positive definite'."""

    def __init__(self, order):
        super().__init__(f"the leading minor of order {order} is not positive definite")
        self.order = order


def _f(a):
    """column-major float64 copy/view"""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _dp(a):
    return a.ctypes.data_as(_lib.c_dp)


def _dpp(arrs):
    arr = (_lib.c_dp * max(1, len(arrs)))()
    for i, a in enumerate(arrs):
        arr[i] = _dp(a)
    return arr


class Context:
    def __init__(self, n_devices=1, device_ids=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        ids = None
        if device_ids is not None:
            ids = (C.c_int * len(device_ids))(*device_ids)
            n_devices = len(device_ids)
        rc = self.lib.fgp_ctx_create(n_devices, ids, C.byref(h))
        if rc != 0:
            raise FgpError(f"fgp_ctx_create failed ({rc}): no sm_100 CUDA device visible -- fungp_b200 has no CPU path")
        self.h = h
        self._problems = weakref.WeakSet()
        # every L_out of this mirror comes from numpy.zeros (calloc): the library need not clear the upper triangle
        self.set_option("out_prezeroed", 1)

    def close(self):
        if getattr(self, "h", None):
            for p in list(self._problems):   # problems hold a pointer to the context: release them first
                p.close()
            self.lib.fgp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, what):
        if rc < 0:
            raise FgpError(f"{what} failed ({rc}): {self.lib.fgp_last_error(self.h).decode()}")
        if rc > 0:
            raise NotPosDefError(rc)

    @property
    def device_count(self):
        return self.lib.fgp_ctx_device_count(self.h)

    def set_option(self, name, value):
        self.check(self.lib.fgp_set_option(self.h, name.encode(), float(value)), "fgp_set_option")

    def timing(self):
        out = np.zeros(8)
        self.check(self.lib.fgp_get_timing(self.h, _dp(out)), "fgp_get_timing")
        return dict(assembly_ms=out[0], factor_ms=out[1], update_ms=out[2], update_launches=int(out[3]),
                    update_flops=out[4], potf2_ms=out[5], trsm_ms=out[6], launches=int(out[7]))

    def last_factor_ms(self):
        """device time of the dataflow-Cholesky launch of the last nll_batch call (0.0 if it ran launch by launch)"""
        ms = C.c_double()
        self.check(self.lib.fgp_last_factor_ms(self.h, C.byref(ms)), "fgp_last_factor_ms")
        return ms.value

    def launch_count(self):
        return int(self.lib.fgp_launch_count())

    def work_count(self, which):
        """0 likelihood factorisations, 1 leave-one-out factorisations, 2 preMats factorisations, 3 candidates scored"""
        return int(self.lib.fgp_work_count(int(which)))

    def mark(self, idx):
        self.check(self.lib.fgp_mark(self.h, idx), "fgp_mark")

    def elapsed_ms(self, i, j):
        ms = C.c_double()
        self.check(self.lib.fgp_elapsed_ms(self.h, i, j, C.byref(ms)), "fgp_elapsed_ms")
        return ms.value

    def measure_peaks(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.check(self.lib.fgp_measure_peaks(self.h, C.byref(a), C.byref(b), C.byref(c)), "fgp_measure_peaks")
        return dict(dmma_tflops=a.value, dfma_tflops=b.value, copy_gbs=c.value)

    def fit_batch(self, sIn, fIn, y, ants, u01, nugget=1e-8, n_starts=1, n_presample=20, dmax=None, ind_vl=None):
        """fgp_fit_batch / fgp_fit_batch_holdout: score candidate structures (rows of `ants`, wire format of formatSol_ACO).
        u01: list with one array per work item (d_c * max(n_presample, n_starts) uniform draws, column-major d_c x n);
        work items = candidates, or (candidate, replicate) pairs, candidate-major, when ind_vl (n_vl x n_rep, 1-based) is given.
        -> dict(fitness, nll, sig2, theta (items x dmax), info)"""
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
        n = y.size
        s = None if sIn is None else _f(np.asarray(sIn, dtype=np.float64).reshape(n, -1))
        ds = 0 if s is None else s.shape[1]
        fs = [_f(f) for f in (fIn or [])]
        df = len(fs)
        ants = np.ascontiguousarray(np.asarray(ants, dtype=np.int32).reshape(-1, ds + 4 * df + 1))
        ncand = ants.shape[0]
        iv = None
        n_rep = 0
        if ind_vl is not None:
            iv = np.asfortranarray(np.asarray(ind_vl, dtype=np.int32).reshape(len(ind_vl), -1))
            n_rep = iv.shape[1]
        nc = ncand * max(1, n_rep)
        off = np.zeros(nc, dtype=np.int64)
        flat = []
        pos = 0
        for c in range(nc):
            off[c] = pos
            uc = np.asarray(u01[c], dtype=np.float64).ravel()
            flat.append(uc); pos += uc.size
        flat = np.ascontiguousarray(np.concatenate(flat)) if flat else np.zeros(1)
        if dmax is None:
            dmax = ds + sum(f.shape[1] for f in fs)
        k = (C.c_int * max(1, df))(*[f.shape[1] for f in fs])
        fit = np.zeros(nc); nll = np.zeros(nc); s2 = np.zeros(nc); th = np.zeros((nc, dmax)); info = np.zeros(nc, dtype=np.int32)
        common = (self.h, n, ds, None if s is None else _dp(s), df, k, _dpp(fs), _dp(y), ncand, ants.ctypes.data_as(_lib.c_ip),
                  float(nugget), int(n_starts), int(n_presample), _dp(flat), off.ctypes.data_as(C.POINTER(C.c_longlong)), int(dmax))
        outs = (_dp(fit), _dp(nll), _dp(s2), _dp(th), info.ctypes.data_as(_lib.c_ip))
        if iv is None:
            rc = self.lib.fgp_fit_batch(*common, *outs)
        else:
            rc = self.lib.fgp_fit_batch_holdout(*common, n_rep, iv.shape[0], iv.ctypes.data_as(_lib.c_ip), *outs)
        self.check(rc, "fgp_fit_batch")
        return dict(fitness=fit, nll=nll, sig2=s2, theta=th, info=info)

    # ---- dimension reduction (R/7_dimRedFunctions.R)
    def bspline_basis(self, k, p):
        B = np.zeros((k, p), order="F")
        self.check(self.lib.fgp_bspline_basis(k, p, _dp(B)), "fgp_bspline_basis")
        return B

    def project(self, f, p, basis_type="B-splines", B=None):
        """-> (basis k x p', J p' x p', coefs n x p');  B given = re-projection with a stored basis."""
        f = _f(f)
        n, k = f.shape
        pp = p if p > 0 else k
        Bo = np.zeros((k, pp), order="F")
        J = np.zeros((pp, pp), order="F")
        co = np.zeros((n, pp), order="F")
        Bi = None if B is None else _f(B)
        rc = self.lib.fgp_project(self.h, _dp(f), n, k, p, BASIS.get(basis_type, 1), None if Bi is None else _dp(Bi),
                                  _dp(Bo), _dp(J), _dp(co))
        self.check(rc, "fgp_project")
        return Bo, J, co


class Problem:
    """Device-resident data set: what setDistMatrix_S / setDistMatrix_F receive (R/7_distanceFunctions.R:3-22)."""

    def __init__(self, ctx, sIn, coefs, J, dis_type, y):
        self.ctx = ctx
        lib = ctx.lib
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
        n = y.size
        s = None if sIn is None else _f(np.asarray(sIn, dtype=np.float64).reshape(n, -1))
        ds = 0 if s is None else s.shape[1]
        coefs = [_f(c) for c in (coefs or [])]
        J = [_f(j) for j in (J or [])]
        df = len(coefs)
        p = (C.c_int * max(1, df))(*[c.shape[1] for c in coefs])
        dt = (C.c_int * max(1, df))(*[DIST[d] if isinstance(d, str) else int(d) for d in (dis_type or [])])
        h = C.c_void_p()
        rc = lib.fgp_problem_create(ctx.h, n, ds, None if s is None else _dp(s), df, p, dt, _dpp(coefs), _dpp(J), _dp(y),
                                    C.byref(h))
        ctx.check(rc, "fgp_problem_create")
        self.h = h
        self.n = n
        self.d = lib.fgp_problem_nls(h)
        self.ds, self.df = ds, df
        self._keep = (s, coefs, J, y)
        ctx._problems.add(self)

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):
                self.ctx.lib.fgp_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_y(self, y):
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
        self.ctx.check(self.ctx.lib.fgp_problem_set_y(self.h, _dp(y)), "fgp_problem_set_y")

    def dist_max(self):
        out = np.zeros(self.d)
        self.ctx.check(self.ctx.lib.fgp_dist_max(self.h, _dp(out)), "fgp_dist_max")
        return out

    def dist_matrix(self, l):
        out = np.zeros((self.n, self.n), order="F")
        self.ctx.check(self.ctx.lib.fgp_dist_materialize(self.h, l, _dp(out)), "fgp_dist_materialize")
        return out

    def assemble(self, ker, nugget, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        out = np.zeros((self.n, self.n), order="F")
        self.ctx.check(self.ctx.lib.fgp_assemble(self.h, KER[ker], float(nugget), _dp(theta), _dp(out)), "fgp_assemble")
        return out

    def nll_batch(self, ker, nugget, thetas, var_known=None):
        """thetas: (d, B) (column b = point b) or (d,) -> (nll[B], sig2[B], info[B])"""
        th = np.asarray(thetas, dtype=np.float64)
        if th.ndim == 1:
            th = th.reshape(-1, 1)
        th = np.asfortranarray(th)
        assert th.shape[0] == self.d, (th.shape, self.d)
        B = th.shape[1]
        nll = np.zeros(B); sig2 = np.zeros(B); info = np.zeros(B, dtype=np.int32)
        vk = None
        if var_known is not None:
            vk = C.c_double(float(var_known))
        rc = self.ctx.lib.fgp_nll_batch(self.h, KER[ker], float(nugget), B, _dp(th), None if vk is None else C.byref(vk),
                                        _dp(nll), _dp(sig2), info.ctypes.data_as(_lib.c_ip))
        self.ctx.check(rc, "fgp_nll_batch")
        return nll, sig2, info

    def premats(self, ker, nugget, sig2, theta, want_L=True):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        L = np.zeros((self.n, self.n), order="F") if want_L else None
        z = np.zeros(self.n)
        rc = self.ctx.lib.fgp_premats(self.h, KER[ker], float(nugget), float(sig2), _dp(theta),
                                      None if L is None else _dp(L), _dp(z))
        self.ctx.check(rc, "fgp_premats")
        return L, z

    def restore(self, ker, nugget, sig2, theta, L, LInvY):
        """re-create the device model from saved slots (preMats$L, preMats$LInvY) without factoring"""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        L = _f(L); z = np.ascontiguousarray(np.asarray(LInvY, dtype=np.float64).ravel())
        assert L.shape == (self.n, self.n) and z.size == self.n
        self.ctx.check(self.ctx.lib.fgp_problem_restore(self.h, KER[ker], float(nugget), float(sig2), _dp(theta), _dp(L), _dp(z)),
                       "fgp_problem_restore")

    def update_var(self, sig2_new, want_L=True):
        L = np.zeros((self.n, self.n), order="F") if want_L else None
        z = np.zeros(self.n)
        self.ctx.check(self.ctx.lib.fgp_update_var(self.h, float(sig2_new), None if L is None else _dp(L), _dp(z)), "fgp_update_var")
        return L, z

    def update_y(self, y_new):
        y = np.ascontiguousarray(np.asarray(y_new, dtype=np.float64).ravel())
        z = np.zeros(self.n)
        self.ctx.check(self.ctx.lib.fgp_update_y(self.h, _dp(y), _dp(z)), "fgp_update_y")
        return z

    def premats_reuse(self, old, want_L=True):
        """model of this (updated) data set with the hyper-parameters of `old`, re-using old's factor for the leading
        points both share -> (L, LInvY, n_reused)"""
        L = np.zeros((self.n, self.n), order="F") if want_L else None
        z = np.zeros(self.n)
        nre = C.c_int()
        self.ctx.check(self.ctx.lib.fgp_premats_reuse(self.h, old.h, None if L is None else _dp(L), _dp(z), C.byref(nre)),
                       "fgp_premats_reuse")
        return L, z, nre.value

    def _new(self, sNew, coefsNew, m):
        s = None if sNew is None else _f(np.asarray(sNew, dtype=np.float64).reshape(m, -1))
        cs = [_f(c) for c in (coefsNew or [])]
        return s, cs

    def predict(self, m, sNew=None, coefsNew=None, detail="light"):
        s, cs = self._new(sNew, coefsNew, m)
        mean = np.zeros(m); sd = np.zeros(m)
        full = detail == "full"
        Ktp = np.zeros((self.n, m), order="F") if full else None
        Kpp = np.zeros((m, m), order="F") if full else None
        rc = self.ctx.lib.fgp_predict(self.h, m, None if s is None else _dp(s), _dpp(cs), 1 if full else 0, _dp(mean),
                                      _dp(sd), None if Ktp is None else _dp(Ktp), None if Kpp is None else _dp(Kpp))
        self.ctx.check(rc, "fgp_predict")
        out = dict(mean=mean, sd=sd)
        if full:
            out["K.tp"] = Ktp
            out["K.pp"] = Kpp
        return out

    def simulate(self, m, Z, sNew=None, coefsNew=None, nugget_sm=0.0):
        s, cs = self._new(sNew, coefsNew, m)
        Z = _f(np.asarray(Z, dtype=np.float64).reshape(m, -1))
        nsim = Z.shape[1]
        sims = np.zeros((nsim, m), order="F")
        mean = np.zeros(m); sd = np.zeros(m)
        rc = self.ctx.lib.fgp_simulate(self.h, m, None if s is None else _dp(s), _dpp(cs), nsim, _dp(Z), float(nugget_sm),
                                       _dp(sims), _dp(mean), _dp(sd))
        self.ctx.check(rc, "fgp_simulate")
        return dict(sims=sims, mean=mean, sd=sd)

    def loocv(self):
        yp = np.zeros(self.n)
        q2 = C.c_double()
        self.ctx.check(self.ctx.lib.fgp_loocv(self.h, _dp(yp), C.byref(q2)), "fgp_loocv")
        return yp, q2.value
