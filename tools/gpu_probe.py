"""GPU probe: measured FP64 peaks + timing of the likelihood path at several sizes / batch shapes.
Writes gpurun_out/probe.json.  Development aid, not part of the product or the tests."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200  # noqa: E402
from fungp_b200 import synth  # noqa: E402

out = {}
ctx = fungp_b200.Context(1)
out["peaks"] = ctx.measure_peaks()
print("peaks", out["peaks"], flush=True)


def run(n, B, streams, lookahead, ker="gauss", reps=3, profile=False, outer=4):
    ctx.set_option("outer", outer)
    ctx.set_option("streams", streams)
    ctx.set_option("lookahead", lookahead)
    ctx.set_option("profile", 1 if profile else 0)
    sIn, fIn, y = synth.make_data(n, n, 4, [100, 100])
    proj = [ctx.project(f, 5, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    theta = 0.25 * 2.0 * P.dist_max()
    th = np.repeat(theta[:, None], B, axis=1) * (1.0 + 0.01 * np.arange(B))[None, :]
    P.nll_batch(ker, 1e-8, th)  # warm
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        nll, s2, info = P.nll_batch(ker, 1e-8, th)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    fl = B * n ** 3 / 3.0
    rec = dict(n=n, B=B, outer=outer, streams=streams, lookahead=lookahead, sec=t, evals_per_s=B / t, chol_tflops=fl / t / 1e12,
               info=int(info.max()), nll0=float(nll[0]))
    if profile:
        rec["timing"] = ctx.timing()
    P.close()
    print(rec, flush=True)
    return rec


recs = []
for n in (1024, 2048, 4096, 8192):
    for outer in (2, 4, 8):
        recs.append(run(n, 1, 1, 1, outer=outer))
    recs.append(run(n, 1, 1, 0, profile=True, outer=8))
    recs.append(run(n, 1, 1, 0, profile=True, outer=4))
for n, B in ((1024, 64), (2048, 20), (8192, 8)):
    for outer in (2, 4, 8):
        recs.append(run(n, B, 4, 0, outer=outer))
out["runs"] = recs
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1, default=float)
