"""C5 predict timing: repeated calls, look-ahead on/off (development aid)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
n, m = 32768, 2048
sIn, fIn, y = synth.make_data(32768, n + m, 2, [200, 200])
proj = [ctx.project(f[:n], 10, "B-splines") for f in fIn]
coefs = [c for (_, _, c) in proj]; J = [j for (_, j, _) in proj]
cnew = [ctx.project(f[n:], 10, "B-splines", B=B)[2] for f, (B, _, _) in zip(fIn, proj)]
P = fungp_b200.Problem(ctx, sIn[:n], coefs, J, ["L2_bygroup"] * 2, y[:n])
theta = 0.5 * 2.0 * P.dist_max()
nll, s2, info = P.nll_batch("matern3_2", 1e-8, theta)
P.premats("matern3_2", 1e-8, s2[0], theta, want_L=False)
for la in (1, 1, 1, 0, 0):
    ctx.set_option("lookahead", la)
    l0 = ctx.launch_count()
    t0 = time.perf_counter(); pr = P.predict(m, sIn[n:], cnew); dt = time.perf_counter() - t0
    print(f"lookahead {la}: predict {dt*1e3:.1f} ms  ({n*n*m/dt/1e12:.2f} TF)  launches {ctx.launch_count()-l0}", flush=True)
for outer in (1, 2, 4, 8):
    ctx.set_option("lookahead", 1); ctx.set_option("outer_single", outer)
    t0 = time.perf_counter(); pr = P.predict(m, sIn[n:], cnew); dt = time.perf_counter() - t0
    print(f"lookahead 1 outer_single {outer}: predict {dt*1e3:.1f} ms  ({n*n*m/dt/1e12:.2f} TF)", flush=True)
