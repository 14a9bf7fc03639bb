"""World-size-2 gloo test of the N>1 host path on CPU: the library's static block partition (equal-cost contiguous
pieces, boundary tuples shared at sub-tile granularity) covers every sub-tile exactly once, is balanced, and the
allreduced per-rank energies -- each rank evaluating ITS sub-tile ranges from the oracle's tiles -- equal the
single-rank total."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from nwchem_b200 import capi, partition, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _range_energy(st, ora, tup, lo, hi):
    """(E1,E2) of the sub-tiles [lo,hi) of one tuple from the oracle's t3 tiles (ccsd_t_dot.F:101-124 on a mask)."""
    t = st.t
    s, d, e1, e2, _ = ora.tuple_tiles(st, tup)
    R = [t.r(int(b)) for b in tup]
    m = partition.sub_tile_mask(R, lo, hi)
    ev = [t.evl_sorted[t.offset[int(b) - 1]: t.offset[int(b) - 1] + t.r(int(b))] for b in tup]
    den = (-ev[0][:, None, None, None, None, None] - ev[1][None, :, None, None, None, None] - ev[2][None, None, :, None, None, None]
           + ev[3][None, None, None, :, None, None] + ev[4][None, None, None, None, :, None] + ev[5][None, None, None, None, None, :])
    import ctypes
    f = ora.lib().ora_ccsd_t_factor(int(t.restricted), *[ctypes.c_long(int(tup[i])) for i in (3, 4, 5, 0, 1, 2)])
    w = np.where(m, f * d / den, 0.0)
    return float(np.sum(w * d)), float(np.sum(w * (d + s))), (e1, e2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"; os.environ["OMP_WAIT_POLICY"] = "passive"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as ora
    st = synth.physical(synth.shape_tiling("h2o_ccpvdz_c2v", tilesize=20))
    tasks = capi.host_task_list(st)                      # the product's own task list (host code of the library)
    ranges = partition.block_partition(st, rank, world)   # the product's own partition (host code of the library)
    e = np.zeros(2)
    items = torch.zeros(len(tasks), dtype=torch.int64)
    for k, (lo, hi) in enumerate(ranges):                 # oracle stands in for the GPU in this CPU test
        if hi > lo:
            a, b, _ = _range_energy(st, ora, [int(x) for x in tasks[k][:6]], int(lo), int(hi))
            e += (a, b)
            items[k] = int(hi - lo)
    tot = partition.allreduce_sum(e)
    dist.all_reduce(items)
    if rank == 0:
        ref = ora.ccsd_t(st)
        full = [int(np.prod([(st.t.r(int(b)) + 3) // 4 for b in tk[:6]])) for tk in tasks]
        out.put((tot, [ref["e1"], ref["e2"]], items.tolist(), full))
    dist.destroy_process_group()


def test_two_rank_block_partition_and_allreduce_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    tot, ref, items, full = q.get(timeout=300)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert items == full                                   # every sub-tile of every task exactly once
    assert abs(tot[0] - ref[0]) <= 1e-12 and abs(tot[1] - ref[1]) <= 1e-12


def test_block_partition_covers_and_balances_h2o10():
    """(H2O)10 shape, 7 590 tasks and a 6-task prefix: the cuts tile the sub-tile space without gaps or overlaps, at
    most one tuple is shared per boundary, and the cost imbalance is below one sub-tile's weight in a million."""
    st = synth.empty_stores(synth.shape_tiling("h2o10_augccpvtz"), intorb=False)
    st.v2_hash = np.zeros(1, np.int64)
    kl = capi.host_task_list(st)
    for ntasks in (0, 6):
        n = len(kl) if ntasks == 0 else ntasks
        full = np.array([np.prod([(st.t.r(int(b)) + 3) // 4 for b in tk[:6]]) for tk in kl[:n]])
        for world in (2, 8):
            rg = [partition.block_partition(st, r, world, 0, ntasks) for r in range(world)]
            cover = sum(r[:, 1] - r[:, 0] for r in rg)
            assert np.array_equal(cover, full)
            for a, b in zip(rg[:-1], rg[1:]):               # contiguous: where rank r stops, rank r+1 starts
                assert np.all((a[:, 1] == b[:, 0]) | (a[:, 1] == a[:, 0]) | (b[:, 1] == b[:, 0]))
            shared = sum(int(np.count_nonzero(r[:, 1] > r[:, 0])) for r in rg) - n
            assert 0 <= shared <= world - 1
            share = np.array([float(np.sum((r[:, 1] - r[:, 0]) / full)) for r in rg])   # tuples are near-equal cost here
            assert share.max() / share.mean() < 1.02
