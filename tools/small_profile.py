"""per-class kernel time of one batched likelihood call at small n (profile mode: per-launch CUDA events)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
import ctypes as C
for n in (512, 1024, 2048):
    sIn, fIn, y = synth.make_data(1024, n, 5, [10, 22, 50])
    ctx = fungp_b200.Context(1)
    proj = [ctx.project(f, 3, "B-splines") for f in fIn[:2]]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    th = 0.4 * 2.0 * P.dist_max()
    for B in (1, 15, 60):
        thb = np.repeat(th[:, None], B, axis=1) * (1 + 0.01 * np.arange(B))[None, :]
        ctx.set_option("profile", 0)
        P.nll_batch("matern5_2", 1e-8, thb)
        t0 = time.perf_counter()
        for _ in range(10): P.nll_batch("matern5_2", 1e-8, thb)
        wall = (time.perf_counter() - t0) / 10
        ctx.set_option("profile", 1)
        P.nll_batch("matern5_2", 1e-8, thb)
        P.nll_batch("matern5_2", 1e-8, thb)
        t = (C.c_double * 8)()
        ctx.lib.fgp_get_timing(ctx.h, t)
        print(f"n={n} B={B}: wall {wall*1e3:.3f} ms | assembly {t[0]:.3f} factor(all) {t[1]:.3f}: update {t[2]:.3f} ({int(t[3])} launches, {t[4]/1e9:.2f} GF -> {t[4]/1e9/max(t[2],1e-9):.1f} TF) potf2 {t[5]:.3f} trsm {t[6]:.3f}  launches {int(t[7])}", flush=True)
    P.close(); ctx.close()
