"""CPU gate for the N > 1 path: world_size-2 gloo run of the host-side sharding / gather logic (the data path has no
device collective; items are independent, so the sharded result must equal the single-rank result bit for bit)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from fungp_b200.shard import shard_range, gather_scalars, sharded_nll


def test_shard_ranges_partition():
    for B in (0, 1, 7, 13, 20, 2000):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


class _FakeProblem:
    """stands in for fungp_b200.Problem on a box without a GPU: a deterministic function of theta"""

    def nll_batch(self, ker, nugget, thetas, var_known=None):
        v = np.sum(np.sin(thetas) * np.arange(1, thetas.shape[0] + 1)[:, None], axis=0) + nugget
        return v, np.abs(v) + 1.0, (v > 2.5).astype(np.int32)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    th = rng.uniform(0.1, 3.0, size=(4, 13))
    nll, s2, info = sharded_nll(_FakeProblem(), "gauss", 1e-8, th, rank, world)
    part = gather_scalars(np.arange(*shard_range(13, rank, world)), 13, rank, world)
    q.put((rank, nll, s2, info, part))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_rank():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    th = rng.uniform(0.1, 3.0, size=(4, 13))
    nll1, s21, info1 = _FakeProblem().nll_batch("gauss", 1e-8, th)
    for rank, nll, s2, info, part in res:
        assert np.array_equal(nll, nll1) and np.array_equal(s2, s21) and np.array_equal(info, info1)
        assert np.array_equal(part, np.arange(13))
