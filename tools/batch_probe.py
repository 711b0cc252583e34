"""headline workload (13 FD points at n=8192) vs outer-panel width / streams (development aid)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
import bench
ctx = fungp_b200.Context(1)
sIn, fIn, y = bench.build_inputs()
proj = [ctx.project(f, 5, "B-splines") for f in fIn]
P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
ub = 2.0 * P.dist_max()
th = synth.fd_points(0.25 * ub, np.full(P.d, 1e-10), ub)
for B in (13, 26):
    thb = np.concatenate([th] * (B // 13), axis=1)
    for outer in (4, 8, 16):
        for streams in (2, 4, 8):
            ctx.set_option("outer", outer); ctx.set_option("streams", streams)
            P.nll_batch("gauss", 1e-8, thb)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); P.nll_batch("gauss", 1e-8, thb); ts.append(time.perf_counter() - t0)
            t = min(ts)
            print(f"B={B} outer={outer} streams={streams}: {t*1e3:.1f} ms  {B/t:.1f} evals/s  {B*8192**3/3/t/1e12:.2f} TF", flush=True)
