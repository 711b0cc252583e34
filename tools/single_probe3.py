"""single-evaluation latency (look-ahead path): adaptive panel widths vs fixed widths, programmatic dependent launch on/off
(development aid).  usage: python tools/single_probe3.py [n ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384]
for n in sizes:
    sIn, fIn, y = synth.make_data(n, n, 4, [100, 100])
    proj = [ctx.project(f, 5, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    theta = 0.25 * 2.0 * P.dist_max()
    ref = None
    for pdl, os_ in ((1, 0), (0, 0), (1, 2), (0, 2), (1, 4), (1, 8)):
        ctx.set_option("pdl", pdl); ctx.set_option("outer_single", os_)
        v = P.nll_batch("gauss", 1e-8, theta)[0][0]
        ref = v if ref is None else ref
        ts = []
        for _ in range(7):
            t0 = time.perf_counter(); P.nll_batch("gauss", 1e-8, theta); ts.append(time.perf_counter() - t0)
        print(f"n={n} pdl={pdl} outer_single={os_ or 'adaptive'}: min {min(ts)*1e3:.3f} ms  median {np.median(ts)*1e3:.3f} ms  "
              f"TF {n**3/3/min(ts)/1e12:.2f}  nll diff vs first {abs(v - ref):.2e}", flush=True)
    ctx.set_option("pdl", 1); ctx.set_option("outer_single", 0)
    P.close()
