"""Where the SM time of the dataflow Cholesky goes (needs a library built with `make EXTRA=-DFGP_DAG_PROF`; development aid).
usage: python tools/dag_phase_probe.py n:B [n:B ...]"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
lib = ctx.lib
rd = lib.fgp_dag_prof_read
rd.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
for a in sys.argv[1:] or ["1024:395", "8192:13"]:
    n, B = (int(v) for v in a.split(":"))
    sIn, fIn, y = synth.make_data(n, n, 4, [100, 100])
    proj = [ctx.project(f, 5, "B-splines") for f in fIn]
    P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
    ub = 2.0 * P.dist_max()
    th = np.stack([0.25 * ub * (1 + 0.0005 * i) for i in range(B)], axis=1)
    P.nll_batch("gauss", 1e-8, th)
    buf = (C.c_ulonglong * 16)()
    rd(buf, 1)
    P.nll_batch("gauss", 1e-8, th)
    ms = ctx.last_factor_ms()
    rd(buf, 1)
    v = [int(x) for x in buf]
    tot = v[6]
    print(f"n={n} B={B}: kernel {ms:.3f} ms, {v[4]} tasks ({v[5]} diagonal); share of CTA time: claim+scan {v[0]/tot:.3f}  tile loop {v[1]/tot:.3f}  "
          f"diag epilogue {v[2]/tot:.3f}  solve epilogue {v[3]/tot:.3f}  unaccounted (start/end idle) {1-(v[0]+v[1]+v[2]+v[3])/tot:.3f};  per task us: "
          f"[solve: hand-over+W wait {v[8]/max(1,v[4]-v[5])/1965:.2f}  products {v[9]/max(1,v[4]-v[5])/1965:.2f}  stores+fence {v[10]/max(1,v[4]-v[5])/1965:.2f}]  "
          f"claim {v[0]/v[4]/1965:.2f}  loop {v[1]/v[4]/1965:.2f}  diag {v[2]/max(1,v[5])/1965:.2f}  solve {v[3]/max(1,v[4]-v[5])/1965:.2f}", flush=True)
    P.close()
