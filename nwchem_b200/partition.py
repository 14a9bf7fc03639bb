"""Host-side view of the static block partition that replaces the reference's shared nxtask counter
(util_gnxtval.c:31, ccsd_t.F:174-255).

The partition itself is computed by the library (csrc/host_driver.h block_partition, the code behind
nwc_triples_run_partition): tasks of the heaviest-first list laid end to end, every 4^6 sub-tile weighted by the k4
planes its tuple contracts, rank r of W takes the r-th equal-cost contiguous piece, boundary tuples shared at sub-tile
granularity.  This module only exposes it without a device (nwc_host_block_partition) and gives the sub-tile order a
numpy meaning, so CPU tests can evaluate a rank's share with the oracle's tiles.
"""
from __future__ import annotations
import ctypes as C
import numpy as np
from . import capi


def block_partition(st, rank: int, world: int, first_task: int = 0, ntasks: int = 0) -> np.ndarray:
    """ranges[i] = (item_lo, item_hi) of task first_task+i for `rank` (host only, no device)."""
    s, keep = capi.make_state(st)
    n_all = len(capi.host_task_list(st))
    n = n_all - first_task if ntasks <= 0 else min(ntasks, n_all - first_task)
    out = np.zeros((max(n, 1), 2), np.int64)
    l = capi.lib()
    l.nwc_host_block_partition.argtypes = [C.POINTER(capi.TceState), C.c_long, C.c_long, C.c_long, C.c_long,
                                           C.POINTER(C.c_longlong)]
    rc = l.nwc_host_block_partition(C.byref(s), rank, world, first_task, ntasks, out.ctypes.data_as(C.POINTER(C.c_longlong)))
    if rc != 0:
        raise RuntimeError("nwc_host_block_partition failed")
    return out[:n]


def sub_tile_mask(ranges_p4p5p6h1h2h3, item_lo: int, item_hi: int) -> np.ndarray:
    """Boolean mask over a t3 tile indexed [p4,p5,p6,h1,h2,h3] of the elements that belong to sub-tiles
    [item_lo, item_hi): sub-tiles are 4-wide blocks numbered with the h3 block fastest and the p4 block slowest
    (TupleHdr.item_first, csrc/kernels.cu)."""
    R = [int(x) for x in ranges_p4p5p6h1h2h3]
    phys = [R[5], R[4], R[3], R[2], R[1], R[0]]            # h3,h2,h1,p6,p5,p4
    nb = [(r + 3) // 4 for r in phys]
    idx = np.indices(R[::-1][::-1], sparse=True)            # per-axis index arrays for [p4,p5,p6,h1,h2,h3]
    blk = [idx[5 - q] // 4 for q in range(6)]               # block coordinate along physical position q
    item = np.zeros(R, np.int64)
    mul = 1
    for q in range(6):
        item = item + blk[q] * mul
        mul *= nb[q]
    return (item >= item_lo) & (item < item_hi)


def allreduce_sum(vec, group=None):
    """Host-side stand-in of nwc_triples_allreduce_energy for CPU (gloo) runs."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(vec), dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t]
