"""Gate 2d (GPU, >= 2 devices): in-process sharding over the GPUs of one context (the path the single-process R
session uses): fgp_nll_batch deals the theta-points to the devices, fgp_fit_batch deals candidates to worker contexts
on every device; per-item results must be bit-identical to the single-GPU run (items are independent)."""
import numpy as np
import pytest

from helpers import make_case, oracle_dists, theta_frac, synth_problem
from oracle.rrng import RRng

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs")
def test_nll_batch_sharded_over_two_gpus_is_bit_identical():
    import fungp_b200
    case = make_case(41, 1500, 3, [40, 40], [4, 3], ["B-splines", "PCA"], ["L2_bygroup", "L2_byindex"])
    Ms = oracle_dists(case)
    theta = theta_frac(Ms, 0.4)
    th = np.stack([theta * (1 + 0.02 * i) for i in range(11)], axis=1)
    out = []
    for nd in (1, 2):
        ctx = fungp_b200.Context(nd)
        assert ctx.device_count == nd
        P = fungp_b200.Problem(ctx, case["sIn"], case["coefs"], case["J"], case["dis"], case["y"])
        out.append(P.nll_batch("matern5_2", 1e-8, th))
        P.close(); ctx.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


@pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs")
def test_fit_batch_over_two_gpus_is_bit_identical():
    import fungp_b200
    from fungp_b200 import factory as F
    n, ds = 300, 3
    sIn, fIn, y, _ = synth_problem(95, n, ds, [12, 20])
    rng = np.random.default_rng(6)
    ants = []
    for c in range(20):
        s_act = [i for i in range(ds) if rng.random() < 0.7] or [0]
        f_act = [j for j in range(2) if rng.random() < 0.6]
        ants.append(F.encode_ant(ds, 2, s_act, f_act, [["L2_bygroup", "L2_byindex"][rng.integers(2)] for _ in f_act],
                                 [int(rng.integers(1, 5)) for _ in f_act], ["B-splines"] * len(f_act),
                                 ["gauss", "matern5_2", "matern3_2"][rng.integers(3)]))
    ants = np.stack(ants)
    res = []
    for nd in (1, 2):
        ctx = fungp_b200.Context(nd)
        res.append(F.score_candidates(ctx, sIn, fIn, y, ants, RRng(11), n_starts=1, n_presample=20))
        ctx.close()
    for key in ("fitness", "nll", "sig2", "theta", "info"):
        assert np.array_equal(res[0][key], res[1][key], equal_nan=True), key
