"""Generate the committed golden fixtures from the reference's own known-answer artefacts.

Run ONCE in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
Outputs (committed):
    tests/golden/xfgpm_models.npz   five stored models of data/precalculated_Xfgpm_objects.rda
                                    (R/8_precalculated_Xfgpm_objects.R:12-19), full double precision
    tests/golden/xfgpm_logs.json    the (structure string -> Q2) pairs of @log.success for each object
    tests/golden/docs_fgpm.json     values printed in docs/reference/fgpm.html (rendered @examples of
                                    R/1_fgpm_Class.R:199-314): nll finals, hypers, B-spline bases,
                                    projection coefficients, 25x25 training covariance
Nothing in tests/ reads /root/reference at run time.
"""
import html
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.rdx3 import read_rda, as_array  # noqa: E402

REF = "/root/reference"


def dump_models():
    objs = read_rda(os.path.join(REF, "data", "precalculated_Xfgpm_objects.rda"))
    out = {}
    logs = {}
    for name, xm in objs.items():
        m = xm.slot("model")
        k = m.slot("kern")
        fp = m.slot("f_proj")
        pm = m.slot("preMats")
        df = int(m.slot("df").value[0])
        ds = int(m.slot("ds").value[0])
        out[f"{name}.ds"] = np.array(ds)
        out[f"{name}.df"] = np.array(df)
        out[f"{name}.kerType"] = np.array(k.slot("kerType").value[0])
        out[f"{name}.nugget"] = np.array(m.slot("nugget").value[0])
        out[f"{name}.negLogLik"] = np.array(m.slot("negLogLik").value[0])
        out[f"{name}.varHyp"] = np.array(k.slot("varHyp").value[0])
        out[f"{name}.s_lsHyps"] = np.asarray(k.slot("s_lsHyps").value, dtype=np.float64)
        out[f"{name}.f_lsHyps"] = np.asarray(k.slot("f_lsHyps").value, dtype=np.float64)
        out[f"{name}.sOut"] = as_array(m.slot("sOut")).ravel()
        out[f"{name}.L"] = as_array(pm["L"])
        out[f"{name}.LInvY"] = as_array(pm["LInvY"]).ravel()
        out[f"{name}.fitness"] = np.array(xm.slot("fitness").value[0])
        out[f"{name}.call"] = np.array(m.slot("howCalled").slot("string").value[0])
        if ds > 0:
            out[f"{name}.sIn"] = as_array(m.slot("sIn"))
        if df > 0:
            out[f"{name}.f_disType"] = np.array(list(k.slot("f_disType").value))
            out[f"{name}.f_pdims"] = np.asarray(fp.slot("pdims").value, dtype=np.int64)
            out[f"{name}.f_basType"] = np.array([("NA" if b is None else b) for b in fp.slot("basType").value])
            for i, (f, B, c) in enumerate(zip(m.slot("fIn").value, fp.slot("basis").value, fp.slot("coefs").value)):
                out[f"{name}.fIn{i}"] = as_array(f)
                out[f"{name}.basis{i}"] = as_array(B)
                out[f"{name}.coefs{i}"] = as_array(c)
        # full data of the factory call (all 5 scalar + 2 functional inputs)
        out[f"{name}.full_sIn"] = as_array(xm.slot("sIn"))
        for i, f in enumerate(xm.slot("fIn").value):
            out[f"{name}.full_fIn{i}"] = as_array(f)
        out[f"{name}.full_sOut"] = as_array(xm.slot("sOut")).ravel()
        ls = xm.slot("log.success")
        logs[name] = dict(calls=[a.slot("string").value[0] for a in ls.slot("args").value],
                          fitness=[float(x) for x in ls.slot("fitness").value])
    np.savez_compressed(os.path.join(HERE, "xfgpm_models.npz"), **out)
    json.dump(logs, open(os.path.join(HERE, "xfgpm_logs.json"), "w"), indent=0)
    print("models:", list(objs), "log pairs:", sum(len(v["fitness"]) for v in logs.values()))


def parse_r_matrix(lines):
    """Parse R's wrapped matrix print ('#>  [1,] 0.1 0.2' blocks with '[,j]' headers)."""
    cols = {}
    cur = None
    for ln in lines:
        ln = ln[2:] if ln.startswith("#>") else ln
        if re.match(r"^\s*(\[,\d+\]\s*)+$", ln):
            cur = [int(x) for x in re.findall(r"\[,(\d+)\]", ln)]
            continue
        mm = re.match(r"^\s*\[(\d+),\]\s+(.*)$", ln)
        if mm and cur:
            r = int(mm.group(1))
            vals = [float(x) for x in mm.group(2).split()]
            for c, v in zip(cur, vals):
                cols.setdefault(c, {})[r] = v
    nr = max(max(d) for d in cols.values())
    nc = max(cols)
    M = np.full((nr, nc), np.nan)
    for c, d in cols.items():
        for r, v in d.items():
            M[r - 1, c - 1] = v
    return M


def block_after(text, marker, stop_markers):
    i = text.index(marker)
    j = min([text.index(s, i + len(marker)) for s in stop_markers if s in text[i + len(marker):]])
    return text[i + len(marker):j].splitlines()


def dump_docs():
    s = open(os.path.join(REF, "docs", "reference", "fgpm.html")).read()
    t = html.unescape(re.sub(r"<[^>]+>", "", s))
    coefs_blk = block_after(t, "m1@f_proj@coefs # list of projection coefficients", ["m1@f_proj@basis"])
    basis_blk = block_after(t, "m1@f_proj@basis # list of projection basis functions", ["Map(function(a, b)"])
    ktt_blk = block_after(t, "tcrossprod(m1@preMats$L) # training auto-covariance matrix", ["# making predictions", "# generating input data for prediction", "n.pr <-"])

    def split2(blk):
        idx = [k for k, l in enumerate(blk) if "[[2]]" in l][0]
        return parse_r_matrix(blk[:idx]), parse_r_matrix(blk[idx:])

    c1, c2 = split2(coefs_blk)
    b1, b2 = split2(basis_blk)
    ktt = parse_r_matrix(ktt_blk)
    doc = dict(
        source="docs/reference/fgpm.html (pkgdown render of the @examples in R/1_fgpm_Class.R:199-314)",
        seed=100, n=25,
        ms=dict(nll=26.743784, var=4.0588, ls=[0.7555, 0.2091]),
        mf=dict(nll=40.291940, var=5.0745, ls=[0.4093, 3.0370]),
        msf=dict(nll=2.841058, var=1.6404, ls=[2.0000, 2.0000, 2.5804, 3.0370]),
        custom=dict(nll=18.777358, var=1.2483,
                    ls_first8=[1.6536, 1.1010, 1.6516, 1.7186, 1.6870, 1.9054, 1.9191, 1.9578]),
        coefs=[c1.tolist(), c2.tolist()], basis=[b1.tolist(), b2.tolist()], Ktt=ktt.tolist(),
        runif_first5=[0.30776611, 0.25767250, 0.55232243, 0.05638315, 0.46854928],
    )
    assert c1.shape == (25, 3) and c2.shape == (25, 3) and b1.shape == (10, 3) and b2.shape == (22, 3), (c1.shape, b1.shape)
    assert ktt.shape == (25, 25) and not np.isnan(ktt).any(), ktt.shape
    json.dump(doc, open(os.path.join(HERE, "docs_fgpm.json"), "w"))
    print("docs: coefs", c1.shape, c2.shape, "basis", b1.shape, b2.shape, "Ktt", ktt.shape)


if __name__ == "__main__":
    dump_models()
    dump_docs()
