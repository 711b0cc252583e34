"""assembly kernel time per correlation family at n = 8192 (C3 structure) and n = 32768 (C5 structure), profile mode"""
import os, sys
import numpy as np
import ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
for n, ds, flen, p in ((8192, 4, 100, 5), (8192, 2, 100, 5), (32768, 2, 200, 10)):
    sIn, fIn, y = synth.make_data(n, n, ds, [flen, flen])
    proj = [ctx.project(f, p, "B-splines") for f in fIn]
    for dis in ("L2_bygroup", "L2_byindex"):
        if dis == "L2_byindex" and n > 8192:
            continue
        P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], [dis] * 2, y)
        theta = 0.3 * 2.0 * P.dist_max()
        for ker in ("gauss", "matern5_2", "matern3_2"):
            ctx.set_option("profile", 1)
            P.nll_batch(ker, 1e-6, theta); P.nll_batch(ker, 1e-6, theta)
            t = (C.c_double * 8)()
            ctx.lib.fgp_get_timing(ctx.h, t)
            ctx.set_option("profile", 0)
            byt = 8.0 * n * (n + 1) / 2
            print(f"n={n} ds={ds} p={p} {dis} {ker}: assembly {t[0]*1e3:.0f} us  {byt/t[0]/1e6:.0f} GB/s ({byt/t[0]/1e6/6542.4*100:.1f} % of HBM copy)  d={P.d}", flush=True)
        P.close()
