"""TMA vs cp.async operand path of the DMMA GEMM inside real factorisations: batched step and single evaluation at n = 8192,
small-n batches.  usage: python tools/gemm_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
for n, B, reps in ((8192, 13, 5), (8192, 1, 7), (2048, 20, 10), (1024, 60, 20)):
    sIn, fIn, y = synth.make_data(n, n, 4, [100, 100])
    proj = [ctx.project(f, 5, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    ub = 2.0 * P.dist_max()
    th = np.stack([0.25 * ub * (1 + 0.001 * i) for i in range(B)], axis=1)
    res = {}
    for tma in (0, 1, 0, 1):
        ctx.set_option("gemm_tma", tma)
        v = P.nll_batch("gauss", 1e-8, th)[0]
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); P.nll_batch("gauss", 1e-8, th); ts.append(time.perf_counter() - t0)
        res.setdefault(tma, []).append(min(ts))
        same = np.array_equal(v, res.setdefault("v", v))
        print(f"n={n} B={B} tma={tma}: min {min(ts)*1e3:8.3f} ms  {B/min(ts):9.1f} evals/s  {B*n**3/3/min(ts)/1e12:6.2f} TFLOP/s  bit-identical {same}", flush=True)
    ctx.set_option("gemm_tma", 1)
    P.close()
