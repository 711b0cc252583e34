"""factory throughput vs number of worker contexts per GPU (development aid)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth, factory as F
from fungp_b200.model import NumpyRng
import bench
ctx = fungp_b200.Context(1)
sIn, fIn, y = synth.make_data(1024, 1024, 5, [10, 22, 50])
ants = bench.factory_candidates(96)
base = None
for w, blk, gph in ((8, 0, 0), (8, 0, 1), (8, 0, 0), (8, 0, 1), (8, 2, 1)):
    ctx.set_option("graphs", gph)
    ctx.set_option("fit_blocking", blk)
    ctx.set_option("fit_workers", w)
    t0 = time.perf_counter()
    r = F.score_candidates(ctx, sIn, fIn, y, ants, NumpyRng(1))
    dt = time.perf_counter() - t0
    bad = np.nonzero(r["info"])[0]
    print(f"workers {w} blocking {blk} graphs {gph}: {len(ants)/dt:.1f} cand/s; failures {[(int(i), int(r['info'][i])) for i in bad]}", flush=True)
    if base is None:
        base = r
    else:
        diff = [int(i) for i in range(len(ants)) if not (np.array_equal(r['nll'][i], base['nll'][i], equal_nan=True))]
        print("   differs from 1-worker run at:", diff[:20], [ (float(r['nll'][i]), float(base['nll'][i])) for i in diff[:4]])
