"""predict / simulate with the dataflow kernel's solve mode against the launch-per-step schedule: bit-identity and time.
usage: python tools/predict_dag_probe.py [n:m ...]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(300, 40), (1000, 10), (1024, 512), (2048, 512), (8192, 64), (8192, 2048)]
for n, m in cases:
    sIn, fIn, y = synth.make_data(n, n + m, 4, [100, 100])
    proj = [ctx.project(f[:n], 5, "B-splines") for f in fIn]
    cnew = [ctx.project(f[n:], 5, "B-splines", B=B)[2] for f, (B, _, _) in zip(fIn, proj)]
    P = fungp_b200.Problem(ctx, sIn[:n], [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y[:n])
    theta = 0.25 * 2.0 * P.dist_max()
    _, s2, _ = P.nll_batch("gauss", 1e-8, theta)
    P.premats("gauss", 1e-8, s2[0], theta, want_L=False)
    Z = np.random.default_rng(1).standard_normal((m, 5))
    ref = {}
    for dag in (0, 1, 0, 1):
        ctx.set_option("dag", dag)
        pr = P.predict(m, sIn[n:], cnew)
        tp = []
        for _ in range(4):
            t0 = time.perf_counter(); P.predict(m, sIn[n:], cnew); tp.append(time.perf_counter() - t0)
        sm = P.simulate(m, Z, sIn[n:], cnew, nugget_sm=1e-8)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); P.simulate(m, Z, sIn[n:], cnew, nugget_sm=1e-8); ts.append(time.perf_counter() - t0)
        tl = []
        for _ in range(3):
            P.premats("gauss", 1e-8, s2[0], theta, want_L=False)          # drops the cached leave-one-out result
            t0 = time.perf_counter(); yp, q2 = P.loocv(); tl.append(time.perf_counter() - t0)
        r = ref.setdefault("r", (pr["mean"], pr["sd"], sm["sims"], yp))
        same = [bool(np.array_equal(a, b)) for a, b in zip(r, (pr["mean"], pr["sd"], sm["sims"], yp))]
        print(f"n={n} m={m} dag={dag}: predict {min(tp)*1e3:8.3f} ms  simulate {min(ts)*1e3:8.3f} ms  loocv {min(tl)*1e3:8.3f} ms  identical(mean, sd, sims, loocv) {same}", flush=True)
    ctx.set_option("dag", 1)
    P.close()
