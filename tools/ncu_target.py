"""Small target for ncu: one likelihood evaluation at n (default 8192), single stream, no look-ahead."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fungp_b200
from fungp_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = fungp_b200.Context(1)
ctx.set_option("lookahead", 0); ctx.set_option("streams", 1)
sIn, fIn, y = synth.make_data(n, n, 4, [100, 100])
proj = [ctx.project(f, 5, "B-splines") for f in fIn]
P = fungp_b200.Problem(ctx, sIn, [c for (_, _, c) in proj], [j for (_, j, _) in proj], ["L2_bygroup"] * 2, y)
theta = 0.25 * 2.0 * P.dist_max()
for _ in range(reps):
    print(P.nll_batch("gauss", 1e-8, theta))
