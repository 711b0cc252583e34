"""Target for ncu: two 13-point likelihood steps of the bench workload (n = 8192) -- each is one chol_dag_kernel launch.
usage: ncu --set full --clock-control none -k regex:chol_dag -s 1 -c 1 python tools/ncu_target_dag.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, fungp_b200
from fungp_b200 import synth
ctx = fungp_b200.Context(1)
ctx.set_option("dag_group", int(os.environ.get("FGP_DAG_GROUP", "1")))      # claim-order experiments
ctx.set_option("dag_order", int(os.environ.get("FGP_DAG_ORDER", "1")))
sIn, fIn, y = bench.build_inputs()
C = bench.CFG
proj = [ctx.project(f, p, C["basis"]) for f, p in zip(fIn, C["pdims"])]
P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], [C["dis"]] * len(fIn), y)
ub = 2.0 * P.dist_max()
th = synth.fd_points(C["theta_frac"] * ub, np.full(P.d, 1e-10), ub)
for _ in range(2):
    print(P.nll_batch(C["ker"], C["nugget"], th)[0][:2])
