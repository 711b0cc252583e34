"""Candidate structures of the ant-colony search (fgpm_factory): wire format, decoding and batched scoring.

The ACO engine itself (R/3_ant_search.R, R/3_ant_admin.R:32-293) stays R -- sequential integer/RNG bookkeeping whose
RNG stream decides the selected structure.  What moves to the GPU is the scoring of a generation's ants
(eval_loocv_ACO, R/3_ant_admin.R:592-648): this module encodes structures in the format of formatSol_ACO
(R/3_ant_admin.R:301-373) and calls fgp_fit_batch.
"""
import re

import numpy as np

from .core import KER, DIST, BASIS

KER_NAMES = {v: k for k, v in KER.items()}


def encode_ant(ds, df, s_active, f_active, f_disType, f_pdims, f_basType, kerType):
    """-> int vector of length ds + 4*df + 1.  s_active / f_active: 0-based indices of the active inputs;
    f_disType / f_pdims / f_basType: one entry per ACTIVE functional input."""
    ant = np.full(ds + 4 * df + 1, -1, dtype=np.int32)
    ant[:ds] = 2
    for i in s_active:
        ant[i] = 1
    it = iter(zip(f_disType, f_pdims, f_basType))
    for j in range(df):
        base = ds + 4 * j
        if j in f_active:
            dt, p, bt = next(it)
            ant[base:base + 4] = [1, DIST[dt], int(p), BASIS.get(bt, 1)]
        else:
            ant[base] = 2
    ant[-1] = KER[kerType]
    return ant


def n_lengthscales(ant, ds, df, k):
    """number of length-scales of a structure (getOwners, R/7_distanceFunctions.R:48-58)"""
    d = int(np.sum(ant[:ds] == 1))
    for j in range(df):
        a = ant[ds + 4 * j: ds + 4 * j + 4]
        if a[0] == 1:
            d += 1 if a[1] == DIST["L2_bygroup"] else (int(a[2]) if a[2] > 0 else int(k[j]))
    return d


def decode_ant(ant, ds, df):
    """-> fgpm() keyword arguments (formatSol_ACO)"""
    s_active = [i for i in range(ds) if ant[i] == 1]
    f_active, dis, pd, bas = [], [], [], []
    inv_d = {v: k for k, v in DIST.items()}; inv_b = {v: k for k, v in BASIS.items()}
    for j in range(df):
        a = ant[ds + 4 * j: ds + 4 * j + 4]
        if a[0] == 1:
            f_active.append(j); dis.append(inv_d[int(a[1])]); pd.append(int(a[2])); bas.append(inv_b.get(int(a[3]), "B-splines"))
    return dict(s_active=s_active, f_active=f_active, f_disType=dis, f_pdims=pd, f_basType=bas, kerType=KER_NAMES[int(ant[-1])])


def _rvec(txt):
    """'c("a", "b")' | '"a"' | 'c(1, 3)' | '3' -> list"""
    txt = txt.strip()
    if txt.startswith("c("):
        txt = txt[2:-1]
    return [t.strip().strip('"') for t in txt.split(",")]


def parse_call_string(call, ds, df):
    """An `fgpm(...)` call string as stored by the reference in Xfgpm@log.success@args (getCall_ACO,
    R/3_ant_admin.R:417-512) -> ant vector.  Vector arguments shorter than the number of active inputs are recycled."""
    def arg(name):
        m = re.search(name + r" = (c\([^)]*\)|\"[^\"]*\"|[^,)]+(?:\[[^\]]*\])?)", call)
        return m.group(1) if m else None
    s_active, f_active = [], []
    m = re.search(r"sIn = sIn(\[,([^\]]*)\])?", call)
    if "sIn = sIn" in call:
        if m and m.group(2):
            sel = m.group(2).replace(",drop=FALSE", "").strip()
            if ":" in sel:
                a, b = sel.split(":"); s_active = list(range(int(a) - 1, int(b)))
            else:
                s_active = [int(t) - 1 for t in _rvec(sel)]
        else:
            s_active = list(range(ds))
    m = re.search(r"fIn = fIn(\[([^\]]*)\])?", call)
    if "fIn = fIn" in call:
        if m and m.group(2):
            sel = m.group(2)
            if ":" in sel:
                a, b = sel.split(":"); f_active = list(range(int(a) - 1, int(b)))
            else:
                f_active = [int(t) - 1 for t in _rvec(sel)]
        else:
            f_active = list(range(df))
    nf = len(f_active)
    rec = lambda v, default: (v * nf)[:nf] if len(v) == 1 else v
    dis = rec(_rvec(arg("f_disType")) if arg("f_disType") else ["L2_bygroup"], None) if nf else []
    pd = rec([int(t) for t in _rvec(arg("f_pdims"))] if arg("f_pdims") else [3], None) if nf else []
    bas = rec(_rvec(arg("f_basType")) if arg("f_basType") else ["B-splines"], None) if nf else []
    ker = _rvec(arg("kerType"))[0] if arg("kerType") else "matern5_2"
    return encode_ant(ds, df, s_active, f_active, dis, pd, bas, ker)


def score_candidates(ctx, sIn, fIn, sOut, ants, rng, nugget=1e-8, n_starts=1, n_presample=20, ind_vl=None):
    """eval_loocv_ACO (ind_vl None) / eval_houtv_ACO (ind_vl = n_vl x n_rep matrix of 1-based validation rows) for a
    whole generation: the presample draws are taken from `rng.runif` in ant (and replicate) order, as the sequential
    reference loop would consume them, then every work item is fitted and scored on the GPU(s).  With hold-out the
    result additionally has `fitness_mean` = mean over the replicates that did not crash (R/3_ant_admin.R:719-721)."""
    ants = np.asarray(ants, dtype=np.int32)
    ds = 0 if sIn is None else np.asarray(sIn).reshape(len(sOut), -1).shape[1]
    df = len(fIn or [])
    k = [np.asarray(f).shape[1] for f in (fIn or [])]
    npre = max(n_presample, n_starts)
    n_rep = 1 if ind_vl is None else np.asarray(ind_vl).reshape(len(ind_vl), -1).shape[1]
    u01 = [np.asarray(rng.runif(n_lengthscales(a, ds, df, k) * npre)) for a in ants for _ in range(n_rep)]
    res = ctx.fit_batch(sIn, fIn, sOut, ants, u01, nugget=nugget, n_starts=n_starts, n_presample=n_presample, ind_vl=ind_vl)
    if ind_vl is not None:
        f = res["fitness"].reshape(len(ants), n_rep)
        with np.errstate(invalid="ignore"):
            res["fitness_mean"] = np.array([np.nanmean(r) if np.any(np.isfinite(r)) else np.nan for r in f])
    return res
