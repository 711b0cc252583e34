import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import fungp_b200
from fungp_b200 import synth, model as M, factory as F
ctx = fungp_b200.Context(1)
n = 2048
sIn, fIn, y = synth.make_data(2048, n, 2, [100, 100])
ant = F.encode_ant(2, 2, [0, 1], [0, 1], ["L2_bygroup"] * 2, [5, 5], ["B-splines"] * 2, "matern5_2")
ref = None
for dag in (0, 1, 2, 0, 1, 2, 0, 1, 2):
    ctx.set_option("dag", dag)
    F.score_candidates(ctx, sIn, fIn, y, ant[None, :], M.NumpyRng(1), n_starts=10, n_presample=20)
    dts = []
    for _ in range(4):
        t0 = time.perf_counter()
        rl = F.score_candidates(ctx, sIn, fIn, y, ant[None, :], M.NumpyRng(1), n_starts=10, n_presample=20)
        dts.append(time.perf_counter() - t0)
    dt = min(dts)
    if ref is None: ref = rl
    print(f"C2 lock-step fit dag={dag}: {dt*1e3:.1f} ms  nll {rl['nll'][0]:.12g}  identical {[bool(np.array_equal(ref[k], rl[k], equal_nan=True)) for k in ('nll','fitness','theta')]}", flush=True)
