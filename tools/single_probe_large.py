"""single-evaluation time at n = 16384 / 32768 vs the outer panel width of the look-ahead path (development aid)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
for n in (16384, 32768):
    sIn, fIn, y = synth.make_data(n, n, 2, [200, 200])
    proj = [ctx.project(f, 10, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    theta = 0.5 * 2.0 * P.dist_max()
    for la, os_ in ((1, 2), (1, 4), (1, 8), (1, 16), (0, 16)):
        ctx.set_option("lookahead", la); ctx.set_option("outer_single", os_); ctx.set_option("outer", 16)
        P.nll_batch("matern3_2", 1e-8, theta)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); P.nll_batch("matern3_2", 1e-8, theta); ts.append(time.perf_counter() - t0)
        print(f"n={n} la={la} outer_single={os_}: min {min(ts)*1e3:.1f} ms  TF {n**3/3/min(ts)/1e12:.2f}", flush=True)
    P.close()
