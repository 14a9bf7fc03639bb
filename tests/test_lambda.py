"""Lambda-CCSD(T): the oracle's restatement of lambda_ccsd_t.F + lambda_ccsd_t_left.F, and the library's
nwc_triples_run_lambda against it.

The reference multiplies the right-hand tile (T3 order) and the left-hand tiles (L3 order, as its TCE_SORTACC_6 calls
deliver them) with one running index although its declarations announce a sort in between.  The CPU tests establish
what the parity target is: (i) at tilesize 1 the literal file and the sorted reading are the same number; (ii) only the
sorted reading is tile-size invariant.  The GPU test then holds the library to the sorted reading at several tilings,
i.e. to the literal reference at tilesize 1 through tile-size invariance."""
import numpy as np
import pytest
from nwchem_b200 import synth, tiling as tl

OCC, VIRT = [2, 1], [3, 2]     # two irreps, 3 occupied / 5 virtual alpha orbitals


def _inputs(ts, restricted=True):
    t = tl.make_tiling(OCC, VIRT, ts, restricted)
    return synth.physical(t, intorb=True), synth.physical_lambda(t)


def test_lambda_oracle_literal_equals_sorted_at_tilesize_1_and_sorted_is_tile_invariant(oracle):
    ref = None
    for ts in (1, 2, 3):
        st, lam = _inputs(ts)
        lit = oracle.lambda_ccsd_t(st, lam, sorted=False)
        srt = oracle.lambda_ccsd_t(st, lam, sorted=True)
        if ts == 1:
            assert lit["e1"] == srt["e1"] and lit["e2"] == srt["e2"]          # ranges of 1: L3 order == T3 order
            ref = srt
            assert abs(ref["e1"]) > 1e-8 and abs(ref["e2"] - ref["e1"]) > 1e-9
        else:
            assert abs(srt["e1"] - ref["e1"]) <= 1e-14 and abs(srt["e2"] - ref["e2"]) <= 1e-14, ts
            assert abs(lit["e1"] - ref["e1"]) > 1e-6 * abs(ref["e1"])         # the literal pairing is not
    st, lam = _inputs(2, restricted=False)                                   # UHF-style tiling: same closed-shell numbers
    srt = oracle.lambda_ccsd_t(st, lam, sorted=True)
    assert abs(srt["e1"] - ref["e1"]) <= 1e-14 and abs(srt["e2"] - ref["e2"]) <= 1e-14


def test_lambda_left_tiles_are_the_t_tiles_of_the_transposed_amplitudes(oracle):
    """The structural fact the library uses: the left-hand contraction tiles are ccsd_t_singles / ccsd_t_doubles applied
    to lambda^T -- with lambda_1 = t1^T, lambda_2 = t2^T and f = 0 the left tiles must equal the (T) tiles (L3 vs T3
    order), tile by tile."""
    t = tl.make_tiling(OCC, VIRT, 2)
    st = synth.physical(t, intorb=True)
    lam = synth.physical_lambda(t)
    # lambda := transposed T amplitudes, block by block
    import dataclasses
    y1 = np.zeros_like(lam.y1); y2 = np.zeros_like(lam.y2)
    n = int(lam.y1_hash[0])
    t1off = {int(st.t1_hash[1 + i]): int(st.t1_hash[1 + int(st.t1_hash[0]) + i]) for i in range(int(st.t1_hash[0]))}
    for i in range(n):
        key, off = int(lam.y1_hash[1 + i]), int(lam.y1_hash[1 + n + i])
        h4b, p1b = key // t.nvab + 1, key % t.nvab + t.noab + 1
        src = t1off[h4b - 1 + t.noab * (p1b - t.noab - 1)]
        blk = st.t1[src:src + t.r(p1b) * t.r(h4b)].reshape(t.r(p1b), t.r(h4b))
        y1[off:off + blk.size] = blk.T.ravel()
    n = int(lam.y2_hash[0])
    t2off = {int(st.t2_hash[1 + i]): int(st.t2_hash[1 + int(st.t2_hash[0]) + i]) for i in range(int(st.t2_hash[0]))}
    for i in range(n):
        key, off = int(lam.y2_hash[1 + i]), int(lam.y2_hash[1 + n + i])
        k = key
        p2b = k % t.nvab + t.noab + 1; k //= t.nvab
        p1b = k % t.nvab + t.noab + 1; k //= t.nvab
        h5b = k % t.noab + 1; k //= t.noab
        h4b = k + 1
        src = t2off[h5b - 1 + t.noab * (h4b - 1 + t.noab * (p2b - t.noab - 1 + t.nvab * (p1b - t.noab - 1)))]
        dims = (t.r(p1b), t.r(p2b), t.r(h4b), t.r(h5b))
        blk = st.t2[src:src + int(np.prod(dims))].reshape(dims)
        y2[off:off + blk.size] = blk.transpose(2, 3, 0, 1).ravel()
    lamT = dataclasses.replace(lam, y1=y1, y2=y2, f1=np.zeros_like(lam.f1))
    checked = 0
    for tup in oracle.task_list(t)[::5]:
        tup = [int(x) for x in tup[:6]]
        s_ref, d_ref = oracle.tuple_tiles(st, tup)[:2]                        # [p4,p5,p6,h1,h2,h3]
        _, _, td, ys, yd = oracle.lambda_tuple(st, lamT, tup)
        assert np.max(np.abs(td - d_ref)) <= 1e-14
        assert np.max(np.abs(yd.transpose(3, 4, 5, 0, 1, 2) - d_ref)) <= 1e-14
        assert np.max(np.abs(ys.transpose(3, 4, 5, 0, 1, 2) - s_ref)) <= 1e-14
        checked += 1
    assert checked > 5


@pytest.mark.gpu
@pytest.mark.parametrize("shape,ts,restricted,intorb", [("small", 2, True, False), ("small", 3, True, True),
                                                         ("small", 2, False, False), ("h2o", 20, True, False)])
def test_lambda_ccsd_t_gpu_matches_oracle(oracle, shape, ts, restricted, intorb):
    """nwc_triples_run_lambda (two-sided tuples through the LAMBDA instantiation of the fused kernel) against the oracle's
    sorted reading, per task and in total; spin-orbital and `2eorb` V2 storage."""
    from nwchem_b200 import capi
    if shape == "small":
        t = tl.make_tiling(OCC, VIRT, ts, restricted)
    else:
        t = synth.shape_tiling("h2o_ccpvdz_c2v", tilesize=ts, restricted=restricted)
    st = synth.physical(t, intorb=True)
    lam = synth.physical_lambda(t)
    ref = oracle.lambda_ccsd_t(st, lam, sorted=True)
    tr = capi.Triples(0)
    if intorb:
        tr.set_state_2eorb(st)
    else:
        import dataclasses
        tr.set_state(dataclasses.replace(st, orb=None))
    tr.set_lambda(lam)
    e1, e2, pt = tr.run_lambda(per_task=True)
    # the oracle lists tuples in the loop order of lambda_ccsd_t.F, the library in heaviest-first order
    order = sorted(range(len(pt)), key=lambda i: tuple(int(x) for x in tr.task_list()[i][:6]))
    got = pt[order]
    tr.close()
    assert abs(e1 - ref["e1"]) <= 1e-12 and abs(e2 - ref["e2"]) <= 1e-12, (e1, ref["e1"], e2, ref["e2"])
    assert np.max(np.abs(got - ref["per_task"])) <= 1e-13
    assert abs(e1 - ref["e1"]) <= 1e-10 * abs(ref["e1"])


@pytest.mark.gpu
def test_lambda_block_partition_sums_to_total(oracle):
    """nwc_triples_run_lambda_partition: the two sums are additive over sub-tiles, so rank pieces (tuples on a boundary
    shared at sub-tile granularity) add up to the single-rank energies, and those match the oracle."""
    from nwchem_b200 import capi
    import dataclasses
    t = synth.shape_tiling("h2o_ccpvdz_c2v")
    st = synth.physical(t, intorb=True)
    lam = synth.physical_lambda(t)
    ref = oracle.lambda_ccsd_t(st, lam, sorted=True)
    tr = capi.Triples(0)
    tr.set_state(dataclasses.replace(st, orb=None))
    tr.set_lambda(lam)
    e1, e2, pt = tr.run_lambda(per_task=True)
    parts = [tr.run_lambda_partition(r, 3, per_task=True) for r in range(3)]
    tr.close()
    assert abs(e1 - ref["e1"]) <= 1e-12 and abs(e2 - ref["e2"]) <= 1e-12
    assert abs(sum(p[0] for p in parts) - e1) <= 1e-14 and abs(sum(p[1] for p in parts) - e2) <= 1e-14
    assert np.max(np.abs(sum(p[2] for p in parts) - pt)) <= 1e-15
