"""Shared test helpers: golden loaders and seeded synthetic inputs."""
import json
import os

import numpy as np

from oracle.rrng import RRng, r_matrix, expand_grid
from oracle import fungp_ref as ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_NAMES = ["xm", "xm25", "xmc", "xmh", "xms"]


def load_models():
    z = np.load(os.path.join(GOLD, "xfgpm_models.npz"), allow_pickle=False)
    out = {}
    for name in MODEL_NAMES:
        g = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}
        df = int(g["df"])
        m = dict(ds=int(g["ds"]), df=df, kerType=str(g["kerType"]), nugget=float(g["nugget"]),
                 negLogLik=float(g["negLogLik"]), varHyp=float(g["varHyp"]), s_ls=g["s_lsHyps"], f_ls=g["f_lsHyps"],
                 sOut=g["sOut"], L=g["L"], LInvY=g["LInvY"], fitness=float(g["fitness"]), call=str(g["call"]),
                 sIn=g.get("sIn"))
        if df:
            m["f_disType"] = [str(x) for x in g["f_disType"]]
            m["f_pdims"] = [int(x) for x in g["f_pdims"]]
            m["f_basType"] = [str(x) for x in g["f_basType"]]
            m["fIn"] = [g[f"fIn{i}"] for i in range(df)]
            m["basis"] = [g[f"basis{i}"] for i in range(df)]
            m["coefs"] = [g[f"coefs{i}"] for i in range(df)]
        m["full_sIn"] = g["full_sIn"]; m["full_fIn"] = [g["full_fIn0"], g["full_fIn1"]]; m["full_sOut"] = g["full_sOut"]
        out[name] = m
    return out


def load_docs():
    return json.load(open(os.path.join(GOLD, "docs_fgpm.json")))


def docs_example_data():
    """set.seed(100) data of R/1_fgpm_Class.R:199-206; returns (sIn, [f1, f2], sOut, rng positioned after)."""
    rng = RRng(100)
    g = np.linspace(0, 1, 5)
    sIn = expand_grid(g, g)
    f1 = r_matrix(rng.runif(25 * 10), ncol=10)
    f2 = r_matrix(rng.runif(25 * 22), ncol=22)
    sOut = ref.fgp_BB3(sIn, [f1, f2])
    return sIn, [f1, f2], sOut, rng


def smooth_curves(rng, n, k, nharm=8):
    """f(t) = sum_j a_j j^-1 sin(j pi t + phi_j), a~U(-1,1), phi~U(0,2pi)."""
    a = 2.0 * r_matrix(rng.runif(n * nharm), ncol=nharm) - 1.0
    ph = 2.0 * np.pi * r_matrix(rng.runif(n * nharm), ncol=nharm)
    t = np.linspace(0, 1, k)
    j = np.arange(1, nharm + 1)
    return np.einsum("nj,njk->nk", a / j, np.sin(j[None, :, None] * np.pi * t[None, None, :] + ph[:, :, None]))


def synth_problem(seed, n, ds, f_lens, extra=0):
    """Seeded synthetic data set: n (+extra new) points, ds scalar inputs, curves of the given lengths."""
    rng = RRng(seed)
    N = n + extra
    sIn = r_matrix(rng.runif(N * ds), ncol=ds) if ds else None
    fIn = [smooth_curves(rng, N, k) for k in f_lens]
    y = np.zeros(N)
    if ds:
        y = y + sIn[:, 0] + (2 * sIn[:, 1] if ds > 1 else 0.0)
    for i, f in enumerate(fIn):
        t = np.linspace(0, 1, f.shape[1])
        y = y + (4 * np.mean(t * f, axis=1) if i == 0 else np.mean(f, axis=1))
    return sIn, fIn, y, rng


# --------------------------------------------------------------------------------------------------
# shared by the GPU parity tests and bench.py: one synthetic case prepared with the oracle's projection
def make_case(seed, n, ds, f_lens, pdims, bas, dis, extra=0):
    """Returns dict with train/new splits: sIn, coefs, J, basis, y, and the oracle's distance lists (train)."""
    sIn, fIn, y, rng = synth_problem(seed, n, ds, f_lens, extra=extra)
    N = n + extra
    case = dict(n=n, m=extra, ds=ds, df=len(f_lens), pdims=list(pdims), bas=list(bas), dis=list(dis))
    case["sIn"] = None if sIn is None else sIn[:n]
    case["sNew"] = None if (sIn is None or not extra) else sIn[n:]
    case["fIn"] = [f[:n] for f in fIn]
    case["fNew"] = [f[n:] for f in fIn]
    case["y"] = y[:n]
    if fIn:
        bcj = ref.dimReduction(case["fIn"], pdims, bas)
        case["basis"], case["coefs"], case["J"] = bcj["basis"], bcj["coefs"], bcj["J"]
        case["coefsNew"] = [ref.project_coefs(B, f)[0] for B, f in zip(bcj["basis"], case["fNew"])] if extra else []
    else:
        case["basis"], case["coefs"], case["J"], case["coefsNew"] = [], [], [], []
    return case


def oracle_dists(case, new=False):
    """(train-train) or (train-new, new-new) distance lists of the oracle."""
    if not new:
        Ms = []
        if case["ds"]:
            Ms += ref.setDistMatrix_S(case["sIn"], case["sIn"])
        if case["df"]:
            Ms += ref.setDistMatrix_F(case["coefs"], case["coefs"], case["J"], case["dis"])
        return Ms
    tp, pp = [], []
    if case["ds"]:
        tp += ref.setDistMatrix_S(case["sIn"], case["sNew"]); pp += ref.setDistMatrix_S(case["sNew"], case["sNew"])
    if case["df"]:
        tp += ref.setDistMatrix_F(case["coefs"], case["coefsNew"], case["J"], case["dis"])
        pp += ref.setDistMatrix_F(case["coefsNew"], case["coefsNew"], case["J"], case["dis"])
    return tp, pp


def theta_frac(Ms, frac):
    """length-scales = frac * upper bound (2 * max distance), SURVEY.md 8d"""
    return frac * ref.setBounds(Ms)[1]
