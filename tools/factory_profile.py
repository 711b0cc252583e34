"""Where a lock-step round spends its device time, measured in place (library event timers around every launch, one lane so
that rounds do not overlap; warm caches, unlike an ncu launch list).  usage: python tools/factory_profile.py [N_CAND]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, fungp_b200
from fungp_b200 import synth
ncand = int(sys.argv[1]) if len(sys.argv) > 1 else 96
F = bench.FACTORY
ctx = fungp_b200.Context(device_ids=[0])
sIn, fIn, y = synth.make_data(F["seed"], F["n"], F["ds"], F["f_lens"])
ants = bench.factory_candidates(ncand, F["list_seed"]); u01 = bench.factory_draws(ants)
ctx.set_option("fit_lanes", 1)
ctx.fit_batch(sIn, fIn, y, ants[:32], u01[:32], nugget=F["nugget"])
for tma in (1, 0):
    ctx.set_option("gemm_tma", tma); ctx.set_option("profile", 1)
    w0 = ctx.work_count(0); t0 = time.perf_counter()
    ctx.fit_batch(sIn, fIn, y, ants, u01, nugget=F["nugget"])
    dt = time.perf_counter() - t0; ev = ctx.work_count(0) - w0
    tm = ctx.timing(); ctx.set_option("profile", 0)
    tot = tm["assembly_ms"] + tm["potf2_ms"] + tm["trsm_ms"] + tm["update_ms"]
    print(json.dumps(dict(gemm_tma=tma, candidates=ncand, seconds=dt, evals=ev, us_per_eval_wall=dt / ev * 1e6,
                          assembly_ms=tm["assembly_ms"], potf2_ms=tm["potf2_ms"], panel_solve_ms=tm["trsm_ms"], update_ms=tm["update_ms"],
                          share=dict(assembly=tm["assembly_ms"] / tot, potf2=tm["potf2_ms"] / tot, panel_solve=tm["trsm_ms"] / tot, update=tm["update_ms"] / tot),
                          update_tflops=tm["update_flops"] / tm["update_ms"] / 1e9, timed_kernel_ms=tot, timed_us_per_eval=tot / ev * 1e3)), flush=True)
ctx.close()
