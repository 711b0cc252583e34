"""C3 predict / simulate timing vs outer panel width and look-ahead (development aid)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
for n, m in ((8192, 2048), (2048, 512), (16384, 2048)):
    sIn, fIn, y = synth.make_data(8192, n + m, 4, [100, 100])
    proj = [ctx.project(f[:n], 5, "B-splines") for f in fIn]
    coefs = [c for (_, _, c) in proj]; J = [j for (_, j, _) in proj]
    cnew = [ctx.project(f[n:], 5, "B-splines", B=B)[2] for f, (B, _, _) in zip(fIn, proj)]
    P = fungp_b200.Problem(ctx, sIn[:n], coefs, J, ["L2_bygroup"] * 2, y[:n])
    theta = 0.25 * 2.0 * P.dist_max()
    nll, s2, info = P.nll_batch("gauss", 1e-8, theta)
    P.premats("gauss", 1e-8, s2[0], theta, want_L=False)
    Z = np.random.default_rng(0).standard_normal((m, 100))
    for la, os_ in ((1, 0), (1, 2), (1, 4), (1, 8), (1, 16), (0, 16)):
        ctx.set_option("lookahead", la); ctx.set_option("outer_single", os_); ctx.set_option("outer", 16)
        P.predict(m, sIn[n:], cnew); P.simulate(m, Z, sIn[n:], cnew, 1e-8)
        tp, ts = [], []
        for _ in range(3):
            t0 = time.perf_counter(); P.predict(m, sIn[n:], cnew); tp.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); P.simulate(m, Z, sIn[n:], cnew, 1e-8); ts.append(time.perf_counter() - t0)
        print(f"n={n} m={m} la={la} outer_single={os_}: predict {min(tp)*1e3:.2f} ms ({n*n*m/min(tp)/1e12:.1f} TF)  simulate {min(ts)*1e3:.2f} ms", flush=True)
    P.close()
