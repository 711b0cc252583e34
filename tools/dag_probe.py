"""Dataflow Cholesky (option "dag") against the launch-per-step engine: bit-identity and time of fgp_nll_batch.
usage: python tools/dag_probe.py [n:B ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
for a in sys.argv[1:]:
    if a.startswith("G"):
        ctx.set_option("dag_group", int(a[1:]))
    if a.startswith("O"):
        ctx.set_option("dag_order", int(a[1:]))
cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:] if not a.startswith(("G", "O"))] or [(256, 1), (256, 3), (1024, 1), (1024, 60), (2048, 20), (8192, 1), (8192, 13)]
for n, B in cases:
    sIn, fIn, y = synth.make_data(n, n, 4, [100, 100])
    proj = [ctx.project(f, 5, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    ub = 2.0 * P.dist_max()
    th = np.stack([0.25 * ub * (1 + 0.001 * i) for i in range(B)], axis=1)
    res = {}
    for dag in (0, 1, 0, 1):
        ctx.set_option("dag", dag)
        v = P.nll_batch("gauss", 1e-8, th)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); P.nll_batch("gauss", 1e-8, th); ts.append(time.perf_counter() - t0)
        same = all(np.array_equal(a, b) for a, b in zip(v, res.setdefault("v", v)))
        print(f"n={n} B={B} dag={dag}: min {min(ts)*1e3:8.3f} ms  {B*n**3/3/min(ts)/1e12:6.2f} TFLOP/s  bit-identical {same}  nll[0]={v[0][0]:.12g} info={v[2].tolist()[:3]}", flush=True)
    ctx.set_option("dag", 0)
    P.close()
