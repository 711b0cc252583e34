"""single-evaluation latency (look-ahead path) at several n, repeated (development aid)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
for n in (1024, 2048, 4096, 8192):
    sIn, fIn, y = synth.make_data(n, n, 4, [100, 100])
    proj = [ctx.project(f, 5, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    theta = 0.25 * 2.0 * P.dist_max()
    for la in (1, 0):
        ctx.set_option("lookahead", la)
        for os_ in (1, 2, 4):
            ctx.set_option("outer_single", os_); ctx.set_option("outer", os_ if la == 0 else 16)
            P.nll_batch("gauss", 1e-8, theta)
            ts = []
            for _ in range(7):
                t0 = time.perf_counter(); P.nll_batch("gauss", 1e-8, theta); ts.append(time.perf_counter() - t0)
            print(f"n={n} la={la} outer={os_}: min {min(ts)*1e3:.3f} ms  median {np.median(ts)*1e3:.3f} ms  max {max(ts)*1e3:.3f}  TF {n**3/3/min(ts)/1e12:.2f}", flush=True)
    P.close()
