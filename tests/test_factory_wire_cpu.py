"""Host logic of the candidate ("ant") wire format (formatSol_ACO, R/3_ant_admin.R:301-373; getCall_ACO :417-512):
runs without a GPU.  The 469 call strings are the reference's own Xfgpm@log.success@args (tests/golden/xfgpm_logs.json)."""
import json
import os

import numpy as np
import pytest

from fungp_b200 import factory as F
from fungp_b200.core import KER, DIST, BASIS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DS, DF, K = 5, 2, [10, 22]          # the fgp_BB7 example of the stored searches: 5 scalar + 2 functional inputs


def _calls():
    logs = json.load(open(os.path.join(GOLD, "xfgpm_logs.json")))
    return [(name, c) for name, v in logs.items() for c in v["calls"]]


def test_every_stored_call_parses_to_a_valid_ant():
    calls = _calls()
    assert len(calls) == 469
    for name, call in calls:
        ant = F.parse_call_string(call, DS, DF)
        assert ant.dtype == np.int32 and ant.shape == (DS + 4 * DF + 1,)
        assert set(ant[:DS]) <= {1, 2} and ant[-1] in KER.values()
        for j in range(DF):
            a = ant[DS + 4 * j: DS + 4 * j + 4]
            if a[0] == 1:
                assert a[1] in DIST.values() and 0 <= a[2] <= K[j] and a[3] in BASIS.values()
            else:
                assert a[0] == 2 and (a[1:] == -1).all()              # R/3_ant_search.R:225-231
        assert (ant[:DS] == 1).any() or any(ant[DS + 4 * j] == 1 for j in range(DF))   # no empty structure is ever logged
        # what the call says is what the ant says
        kw = F.decode_ant(ant, DS, DF)
        assert ('kerType = "%s"' % kw["kerType"]) in call or ("kerType" not in call and kw["kerType"] == "matern5_2")
        assert ("sIn = sIn" in call) == bool(kw["s_active"]) and ("fIn = fIn" in call) == bool(kw["f_active"])


def test_encode_decode_round_trip_and_length_scale_count():
    rng = np.random.default_rng(0)
    kers, diss, bass = list(KER), list(DIST), list(BASIS)
    for _ in range(500):
        ds, df = int(rng.integers(0, 6)), int(rng.integers(0, 4))
        k = [int(rng.integers(3, 40)) for _ in range(df)]
        s_act = sorted(rng.choice(ds, size=int(rng.integers(0, ds + 1)), replace=False).tolist()) if ds else []
        f_act = sorted(rng.choice(df, size=int(rng.integers(0, df + 1)), replace=False).tolist()) if df else []
        dis = [diss[int(rng.integers(len(diss)))] for _ in f_act]
        pd = [int(rng.integers(0, k[j] + 1)) for j in f_act]
        bas = [bass[int(rng.integers(len(bass)))] for _ in f_act]
        ker = kers[int(rng.integers(len(kers)))]
        ant = F.encode_ant(ds, df, s_act, f_act, dis, pd, bas, ker)
        kw = F.decode_ant(ant, ds, df)
        assert kw == dict(s_active=s_act, f_active=f_act, f_disType=dis, f_pdims=pd, f_basType=bas, kerType=ker)
        want = len(s_act) + sum(1 if d == "L2_bygroup" else (p if p > 0 else k[j]) for j, d, p in zip(f_act, dis, pd))
        assert F.n_lengthscales(ant, ds, df, k) == want               # getOwners, R/7_distanceFunctions.R:48-58


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_round_robin_deal_covers_every_candidate_once(world):
    """bench.py / an R par.clust driver deal candidates ants[rank::world]: a partition, sizes within one of each other."""
    ants = np.arange(37 * 14).reshape(37, 14)
    parts = [ants[r::world] for r in range(world)]
    assert sum(len(p) for p in parts) == len(ants)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    back = np.empty_like(ants)
    for r, p in enumerate(parts):
        back[r::world] = p
    assert np.array_equal(back, ants)
