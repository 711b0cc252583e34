"""Factory arm of bench.py under different driver options (GPU box): candidates/s, likelihood evaluations/s, TFLOP/s.
usage: python tools/factory_probe2.py N_CAND  [mode:points:lanes[:fused_assembly[:gemm_flat[:dag]]] ...]   e.g.  512 1:512:2 1:256:2 1:1024:1 0:0:0"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import fungp_b200  # noqa: E402
from fungp_b200 import synth  # noqa: E402

ncand = int(sys.argv[1]) if len(sys.argv) > 1 else 256
configs = sys.argv[2:] or ["1:512:2", "0:0:0"]
F = bench.FACTORY
ctx = fungp_b200.Context(device_ids=[0])
sIn, fIn, y = synth.make_data(F["seed"], F["n"], F["ds"], F["f_lens"])
ants = bench.factory_candidates(ncand, F["list_seed"])
u01 = bench.factory_draws(ants)
ref = None
for cfg in configs:
    parts = [int(v) for v in cfg.split(":")]
    mode, pts, lanes = parts[:3]
    fusedA = parts[3] if len(parts) > 3 else 1
    ctx.set_option("gemm_flat", parts[4] if len(parts) > 4 else 8)
    ctx.set_option("dag", parts[5] if len(parts) > 5 else 0)
    ctx.set_option("fit_mode", mode); ctx.set_option("fit_points", pts); ctx.set_option("fit_lanes", lanes)
    ctx.set_option("fused_assembly", fusedA)
    if mode == 0:
        ctx.set_option("lookahead", 0)
    ctx.fit_batch(sIn, fIn, y, ants[:32], u01[:32], nugget=F["nugget"])
    w0 = [ctx.work_count(i) for i in range(4)]; l0 = ctx.launch_count()
    t0 = time.perf_counter()
    res = ctx.fit_batch(sIn, fIn, y, ants, u01, nugget=F["nugget"], n_starts=F["n_starts"], n_presample=F["n_presample"])
    dt = time.perf_counter() - t0
    w1 = [ctx.work_count(i) for i in range(4)]
    ev, lv = w1[0] - w0[0], w1[1] - w0[1]
    tf = (ev * F["n"] ** 3 / 3 + lv * 2 * F["n"] ** 3 / 3) / dt / 1e12
    same = None
    if ref is None:
        ref = res
    else:
        same = all(np.array_equal(ref[k], res[k], equal_nan=True) for k in ("fitness", "nll", "sig2", "theta", "info"))
    print(json.dumps(dict(cfg=cfg, cands_per_s=ncand / dt, seconds=dt, evals=ev, evals_per_cand=ev / ncand, evals_per_s=ev / dt, loocv=lv,
                          tflops=tf, launches=ctx.launch_count() - l0, ok=int((res["info"] == 0).sum()), identical_to_first=same)), flush=True)
    ctx.set_option("lookahead", 1)
ctx.close()
