"""Is the small-n batched regime GPU-bound or launch-bound?  calls/s of 8 concurrent contexts at n = 256..1024."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
from fungp_b200 import _lib
lib = _lib.load()
B = 15
for n in (256, 512, 1024):
    sIn, fIn, y = synth.make_data(1024, n, 5, [10, 22, 50])
    def setup():
        ctx = fungp_b200.Context(1)
        proj = [ctx.project(f, 3, "B-splines") for f in fIn[:2]]
        P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
        th = 0.4 * 2.0 * P.dist_max()
        return ctx, P, th
    for nt in (1, 8):
        ctxs = [setup() for _ in range(nt)]
        def work(i, reps):
            c, p, t = ctxs[i]
            thb = np.repeat(t[:, None], B, axis=1) * (1 + 0.01 * np.arange(B))[None, :]
            for _ in range(reps): p.nll_batch("matern5_2", 1e-8, thb)
        for i in range(nt): work(i, 3)
        l0 = lib.fgp_launch_count()
        ths = [threading.Thread(target=work, args=(i, 40)) for i in range(nt)]
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        dt = time.perf_counter() - t0
        nl = lib.fgp_launch_count() - l0
        print(f"n={n} contexts={nt}: {nt*40/dt:.0f} calls/s, {nt*40*B/dt:.0f} evals/s, {nl/(nt*40):.0f} launches/call, {nl/dt/1e3:.0f}k launches/s, "
              f"{nt*40*B/dt * n**3/3/1e12:.2f} TF", flush=True)
        for c, p, t in ctxs: p.close(); c.close()
