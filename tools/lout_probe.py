"""D2H of the stored factor (preMats$L) with and without the transparent-huge-page hint on the destination.
usage: python tools/lout_probe.py [n ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
for n in [int(v) for v in sys.argv[1:]] or [8192, 32768]:
    sIn, fIn, y = synth.make_data(n, n, 2, [100, 100])
    proj = [ctx.project(f, 5, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    theta = 0.5 * 2.0 * P.dist_max()
    nll, s2, info = P.nll_batch("matern3_2", 1e-8, theta)
    for rep in range(2):
        for thp, nth, pz in ((1, 0, 0), (1, 4, 0), (1, 12, 0), (1, 16, 0), (1, 0, 1), (1, 4, 1), (1, 16, 1), (0, 0, 1)):
            ctx.set_option("host_thp", thp); ctx.set_option("host_threads", nth); ctx.set_option("out_prezeroed", pz)
            t0 = time.perf_counter(); P.premats("matern3_2", 1e-8, s2[0], theta, want_L=False); t_pm = time.perf_counter() - t0
            t0 = time.perf_counter(); L, z = P.premats("matern3_2", 1e-8, s2[0], theta, want_L=True); t_l = time.perf_counter() - t0
            chk = float(L[n - 1, n - 1])
            del L
            print(f"n={n} thp={thp} threads={nth} prezeroed={pz}: premats {t_pm*1e3:8.1f} ms, with L_out {t_l*1e3:8.1f} ms -> L_out {8.0*n*n/(t_l-t_pm)/1e9:6.1f} GB/s  (L[n-1,n-1]={chk:.6g})", flush=True)
    P.close()
