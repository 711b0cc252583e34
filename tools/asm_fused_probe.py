"""assembly kernel time per matrix: per-matrix kernel vs batch-fused kernel (library event timers, profile mode).
usage: python tools/asm_fused_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
for n, ds, B in ((8192, 4, 13), (8192, 2, 9), (2048, 2, 20), (1024, 4, 60)):
    sIn, fIn, y = synth.make_data(n, n, ds, [100, 100])
    proj = [ctx.project(f, 5, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    ub = 2.0 * P.dist_max()
    th = np.stack([0.25 * ub * (1 + 0.001 * i) for i in range(B)], axis=1)
    byt = 8.0 * n * (n + 1) / 2
    for ker in ("gauss", "matern5_2", "matern3_2"):
        row = []
        for fused in (0, 1):
            ctx.set_option("fused_assembly", fused); ctx.set_option("profile", 1)
            P.nll_batch(ker, 1e-8, th); P.nll_batch(ker, 1e-8, th)
            ms = ctx.timing()["assembly_ms"] / B
            ctx.set_option("profile", 0)
            row.append(f"{'fused' if fused else 'plain'} {ms*1e3:8.1f} us/matrix {byt/ms/1e6:7.0f} GB/s ({byt/ms/1e6/6542.4*100:4.1f} % of HBM copy)")
        print(f"n={n} d={P.d} B={B} {ker:10s}: " + " | ".join(row), flush=True)
    ctx.set_option("fused_assembly", 1)
    P.close()
