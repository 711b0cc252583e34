"""Host-side sharding of independent work items (length-scale vectors, candidate structures) over processes / GPUs.

The path has no device collective: every rank evaluates a contiguous share of the batch on its own
GPU and the B scalars per quantity are gathered on the host.  Inside one process libfungp_b200.so does the same over
the GPUs of its context (fgp_nll_batch / fgp_fit_batch, same split formula); this module is the one-process-per-GPU
variant used under torchrun (bench.py) -- the analogue of the reference's foreach/%dorng% over a PSOCK cluster
(R/3_training_SF.R:161-195, R/3_ant_admin.R:596-622).
"""
import numpy as np


def shard_range(n_items, rank, world):
    """[b0, b1) of rank `rank`: b0 = floor(n_items * rank / world) -- identical to the split inside fgp_nll_batch."""
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def gather_scalars(local, n_items, rank, world, group=None):
    """All ranks contribute their share (values for items shard_range(...)); every rank gets the full length-n_items array.
    Results are bit-identical to a single-rank evaluation because items are independent."""
    local = np.asarray(local)
    b0, b1 = shard_range(n_items, rank, world)
    assert local.shape[0] == b1 - b0, (local.shape, b0, b1)
    if world == 1:
        return local.copy()
    import torch
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    return np.concatenate([np.asarray(p) for p in parts], axis=0)


def sharded_nll(problem, ker, nugget, thetas, rank, world, var_known=None):
    """negLogLik for the columns of `thetas` (d x B), this rank evaluating only its share; returns full-length arrays."""
    B = thetas.shape[1]
    b0, b1 = shard_range(B, rank, world)
    if b1 > b0:
        nll, s2, info = problem.nll_batch(ker, nugget, thetas[:, b0:b1], var_known=var_known)
    else:
        nll, s2, info = np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int32)
    return (gather_scalars(nll, B, rank, world), gather_scalars(s2, B, rank, world), gather_scalars(info, B, rank, world))
