"""sweep of the adaptive look-ahead thresholds for a single factorisation (development aid)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
for n in [int(a) for a in sys.argv[1:]] or [8192]:
    sIn, fIn, y = synth.make_data(n, n, 4, [100, 100])
    proj = [ctx.project(f, 5, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    theta = 0.25 * 2.0 * P.dist_max()
    for r8, r4 in ((44, 28), (52, 36), (36, 20), (60, 28), (44, 16), (36, 28), (28, 12), (52, 20), (1000, 28), (1000, 1000)):
        ctx.set_option("la_r8", r8); ctx.set_option("la_r4", r4)
        P.nll_batch("gauss", 1e-8, theta)
        ts = []
        for _ in range(7):
            t0 = time.perf_counter(); P.nll_batch("gauss", 1e-8, theta); ts.append(time.perf_counter() - t0)
        print(f"n={n} la_r8={r8} la_r4={r4}: min {min(ts)*1e3:.3f} ms  median {np.median(ts)*1e3:.3f} ms  TF {n**3/3/min(ts)/1e12:.2f}", flush=True)
    P.close()
