"""n=1024 batched-call latency and multi-context concurrency (development aid)"""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", sys.argv[1] if len(sys.argv) > 1 else "8")
import fungp_b200
from fungp_b200 import synth
n = 1024
sIn, fIn, y = synth.make_data(1024, n, 5, [10, 22, 50])
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
def setup():
    ctx = fungp_b200.Context(1)
    ctx.set_option("streams", NS)
    proj = [ctx.project(f, 3, "B-splines") for f in fIn[:2]]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    th = 0.4 * 2.0 * P.dist_max()
    return ctx, P, th
ctx, P, th = setup()
print('connections', os.environ['CUDA_DEVICE_MAX_CONNECTIONS'], 'streams', NS)
for B in (15, 30, 64):
    thb = np.repeat(th[:, None], B, axis=1) * (1 + 0.01 * np.arange(B))[None, :]
    P.nll_batch("matern5_2", 1e-8, thb)
    t0 = time.perf_counter()
    for _ in range(20): P.nll_batch("matern5_2", 1e-8, thb)
    dt = (time.perf_counter() - t0) / 20
    print(f"B={B}: {dt*1e3:.3f} ms per call, {B/dt:.0f} evals/s", flush=True)
for nt in (2, 4, 8):
    ctxs = [setup() for _ in range(nt)]
    B = 15
    def work(i):
        c, p, t = ctxs[i]
        thb = np.repeat(t[:, None], B, axis=1) * (1 + 0.01 * np.arange(B))[None, :]
        for _ in range(30): p.nll_batch("matern5_2", 1e-8, thb)
    for i in range(nt): work(i)
    ths = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    dt = time.perf_counter() - t0
    print(f"{nt} contexts x 30 calls of B=15: {dt*1e3:.1f} ms -> {nt*30/dt:.0f} calls/s, {nt*30*B/dt:.0f} evals/s", flush=True)
