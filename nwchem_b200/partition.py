"""Static task partition that replaces the reference's shared nxtask counter (util_gnxtval.c:31, ccsd_t.F:174-255).

Rank r of W runs tasks r, r+W, r+2W, ... of the heaviest-first list (ccsd_t_neword.F): exactly the
(first, stride) pair passed to nwc_triples_run.  The two energies are then summed over ranks (ga_dgop,
ccsd_t.F:297): NCCL inside the library on GPUs, torch.distributed (gloo) in the CPU tests.
"""
from __future__ import annotations
import numpy as np


def rank_tasks(ntasks: int, rank: int, world: int) -> range:
    return range(rank, ntasks, world)


def first_stride(rank: int, world: int) -> tuple[int, int]:
    return rank, world


def weights_per_rank(weights, world: int):
    """Sum of task weights each rank receives under the round-robin deal of a heaviest-first list."""
    w = np.asarray(weights, dtype=np.float64)
    return np.array([w[r::world].sum() for r in range(world)])


def allreduce_sum(vec, group=None):
    """Host-side stand-in of nwc_triples_allreduce_energy for CPU (gloo) runs."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(vec), dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t]
