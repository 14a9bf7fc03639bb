"""Randomised tilings (1, 2 or 4 irreps, empty irreps, tile sizes 1..4, restricted and unrestricted, physical and iid
block stores) through the library's host driver on the CPU (trace context) against the oracle's tiles: (T), the
Lambda-CCSD(T) tuple, the one-pass CR-CCSD(T) tuple, the one-tuple CR-EOMCCSD(T) form, and the block partition (the
ranks' sub-tile ranges tile every task exactly once).  A fixed seed and a bounded number of cases; the same loop was run
unbounded for minutes with other seeds (about 3 000 tilings, no mismatch) when it was written."""
import dataclasses
import numpy as np
import pytest
from nwchem_b200 import capi, synth, tiling as tl, partition
from test_trace import evaluate, _energies, _eps_of


def _tilings(rng, n, max_occ, max_virt):
    out = []
    while len(out) < n:
        nirr = int(rng.choice([1, 2, 4]))
        occ = [int(x) for x in rng.integers(0, 4, nirr)]
        virt = [int(x) for x in rng.integers(0, 5, nirr)]
        if sum(occ) < 2 or sum(virt) < 2 or sum(occ) > max_occ or sum(virt) > max_virt:
            continue
        out.append(tl.make_tiling(occ, virt, int(rng.integers(1, 5)), bool(rng.integers(0, 2))))
    return out


def test_random_tilings_t_lambda_cr(oracle):
    from oracle import cr_dense
    rng = np.random.default_rng(20261017)
    ntup = 0
    for t in _tilings(rng, 40, 6, 8):
        tasks = oracle.task_list(t)
        if len(tasks) == 0:
            continue
        physical = bool(rng.integers(0, 2))
        if physical:
            st = synth.physical(t, intorb=True)
            cr = cr_dense.Dense(t).stores()
            lam = synth.physical_lambda(t)
        else:
            st = synth.random_blocks(t, seed=int(rng.integers(1, 100)))
            (n1h, n1), (n2h, n2), (e2h, e2) = tl.cr_n1_offset(t), tl.cr_n2_offset(t), tl.cr_e2_offset(t)
            cr = cr_dense.CRStores(n1h, rng.uniform(-1, 1, n1) * 0.1, n2h, rng.uniform(-1, 1, n2) * 0.1, e2h,
                                   rng.uniform(-1, 1, e2) * 0.02, 0.0)
            lam = None
        tr = capi.Triples(trace=True)
        tr.set_state(dataclasses.replace(st, orb=None) if physical else st)
        tr.set_cr(cr)
        if lam is not None:
            tr.set_lambda(lam)
        for tup in tasks[rng.permutation(len(tasks))[:4]]:
            tup = [int(x) for x in tup[:6]]
            s_ref, d_ref = oracle.tuple_tiles(st, tup)[:2]
            d, _, s, f, two = evaluate(tr.trace_tuple(tup, 0)[0])
            assert not two and np.max(np.abs(d - d_ref)) <= 1e-13 and np.max(np.abs(s - s_ref)) <= 1e-13, (t.range, tup)
            _, m_ref, e_ref = oracle.cr_tuple(st, cr, tup)
            m, d1, s1, f1, two, e = evaluate(tr.trace_tuple(tup, 4)[0])
            assert np.max(np.abs(m - m_ref)) <= 1e-13 and np.max(np.abs(e - e_ref)) <= 1e-13, (t.range, tup)
            assert np.array_equal(d1, d) and np.array_equal(s1, s) and f1 == f
            if lam is not None:
                td, yd, ys, f, two = evaluate(tr.trace_tuple(tup, 1)[0])
                _, _, td_ref, ys_ref, yd_ref = oracle.lambda_tuple(st, lam, tup, sorted=True)
                assert two and np.max(np.abs(td - td_ref)) <= 1e-13, (t.range, tup)
                assert np.max(np.abs(yd - yd_ref.transpose(3, 4, 5, 0, 1, 2))) <= 1e-13
                assert np.max(np.abs(ys - ys_ref.transpose(3, 4, 5, 0, 1, 2))) <= 1e-13
            ntup += 1
        tr.close()
    assert ntup >= 60


def test_random_tilings_creom_and_block_partition(oracle):
    from oracle import cr_dense
    rng = np.random.default_rng(7)
    ntup = 0
    for t in _tilings(rng, 30, 5, 7):
        tasks = oracle.task_list(t)
        if len(tasks) == 0:
            continue
        r0 = float(rng.choice([0.0, 0.37, -1.2]))
        st = synth.physical(t)
        cr, q = cr_dense.DenseEOM(t, r0=r0).stores()
        tr = capi.Triples(trace=True)
        tr.set_state(st)
        if abs(r0) >= 1e-7:
            tr.set_cr(cr)
        tr.set_creom(q)
        for tup in tasks[rng.permutation(len(tasks))[:4]]:
            tup = [int(x) for x in tup[:6]]
            sums_ref, r_ref, l_ref = oracle.cr_eom_tuple(st, cr, q, tup)
            recs = tr.trace_tuple(tup, 8)[0]
            _, r8, l8, f, two = evaluate(recs)
            eps = _eps_of(recs[-1], r8.shape)
            a, apc = _energies(t, tup, r8, r8, l8, f, eps)
            b, bpd = f * np.sum(l8 * r8), f * np.sum(l8 * (r8 + l8))
            got = np.array([a, b, apc - a, bpd - b])
            assert np.max(np.abs(r8 - r_ref)) <= 1e-13 and np.max(np.abs(l8 - l_ref)) <= 1e-13, (t.range, tup)
            assert np.max(np.abs(got - sums_ref)) <= 1e-12 * max(1.0, np.max(np.abs(sums_ref))), (t.range, tup)
            ntup += 1
        tr.close()
        world = int(rng.integers(1, 6))
        parts = [partition.block_partition(st, r, world) for r in range(world)]
        full = partition.block_partition(st, 0, 1)
        for i in range(len(full)):
            cur = int(full[i, 0])
            for a, b in sorted((int(p[i, 0]), int(p[i, 1])) for p in parts if p[i, 1] > p[i, 0]):
                assert a == cur, (t.range, world, i)
                cur = b
            assert cur == int(full[i, 1]), (t.range, world, i)
    assert ntup >= 50


def test_random_tilings_2eorb_host_plan():
    """the `2eorb` plan (which orbital-form block, strides, sign per half) rebuilds every stored spin-orbital V2 block bit
    for bit on random tilings (tests/test_host.py does this on the H2O C2v table)"""
    rng = np.random.default_rng(11)
    nblk = 0
    for t in _tilings(rng, 25, 6, 8):
        st = synth.physical(t, intorb=True)
        vo = st.orb.v2orb
        for key, off in synth._iter_hash(st.v2_hash):
            g3b, g4b, g1b, g2b = tl.decode_v2_key(t, key)
            dims = [t.r(g3b), t.r(g4b), t.r(g1b), t.r(g2b)]
            oa, ob, strides = capi.host_2eorb_plan(st, g3b, g4b, g1b, g2b)
            idx = np.indices(dims).reshape(4, -1)
            blk = np.zeros(idx.shape[1])
            if oa >= 0:
                blk += vo[oa + (idx * strides[0][:, None]).sum(0)]
            if ob >= 0:
                blk -= vo[ob + (idx * strides[1][:, None]).sum(0)]
            assert np.array_equal(blk, st.v2[off:off + int(np.prod(dims))]), (t.range, g3b, g4b, g1b, g2b)
            nblk += 1
    assert nblk >= 1000
