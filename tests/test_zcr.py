"""CR-CCSD(T) (src/tce/ccsd_t/cr_ccsd_t.F): the oracle's restatement of the tuple loop and the library's
nwc_triples_run_cr against it.

The tuple loop needs four t3-sized tiles per tuple -- the (T) singles S and doubles D, the moment M (cr_ccsd_t_N_1/_N_2:
the (T) doubles contractions with V2 replaced by two dressed intermediates) and the denominator tile E (cr_ccsd_t_E_1/_E_2:
outer products of t1 with t2 and with a t1*t1 intermediate) -- and forms four sums, M.D, M.(S+D), E.D, E.(S+D), each over
f/Delta.  The intermediates are inputs of the loop (the reference builds them once, or reads them from files); the tests
take them from a dense spin-orbital evaluation of the TCE expressions (oracle/cr_dense.py).

CPU tests pin the restatement: (i) the tiled sums equal an untiled dense evaluation of the same algebra to 1e-16;
(ii) tile-size invariance and restricted == unrestricted; (iii) M -> D as the amplitudes go to zero (the dressed
intermediates reduce to V2), which ties the conventions of the dense intermediates to the line-by-line (T) oracle.
(File name: sorts last, so the driver's `pytest -x` reaches every older GPU test first.)"""
import dataclasses
import numpy as np
import pytest
from nwchem_b200 import synth, tiling as tl

OCC, VIRT = [2, 1], [3, 2]     # two irreps, 3 occupied / 5 virtual alpha orbitals


def _inputs(ts, restricted=True, shape=None):
    from oracle import cr_dense
    if shape is None:
        t = tl.make_tiling(OCC, VIRT, ts, restricted)
    else:
        t = synth.shape_tiling(shape, tilesize=ts, restricted=restricted)
    d = cr_dense.Dense(t)
    return synth.physical(t, intorb=True), d.stores(), d


def test_cr_offset_tables_have_the_block_structure_of_the_v2_classes_they_dress():
    """OFFSET_cr_ccsd_t_N_1_1 / _N_2_1 enumerate the same tile quadruples as the <hp||hh> / <pp||hp> classes of V2
    (the intermediates start as copies of those blocks, cr_ccsd_t_N.F:669,:3907), in their own key order."""
    t = synth.shape_tiling("h2o_ccpvdz_c2v")
    n1h, n1 = tl.cr_n1_offset(t); n2h, n2 = tl.cr_n2_offset(t); e2h, e2 = tl.cr_e2_offset(t)
    v2h, _ = tl.v2_offset(t)
    hphh = set(); pphp = set()
    for i in range(int(v2h[0])):
        g3, g4, g1, g2 = tl.decode_v2_key(t, int(v2h[1 + i]))
        cls = tuple(b > t.noab for b in (g3, g4, g1, g2))
        if cls == (False, True, False, False): hphh.add((g3, g4, g1, g2))
        if cls == (True, True, False, True): pphp.add((g3, g4, g1, g2))
    got1 = set()
    for i in range(int(n1h[0])):
        p4b, h11b, h1b, h2b = tl.decode_cr_n1_key(t, int(n1h[1 + i]))
        got1.add((h11b, p4b, h1b, h2b))
    got2 = {tl.decode_cr_n2_key(t, int(n2h[1 + i])) for i in range(int(n2h[0]))}
    assert got1 == hphh and got2 == pphp
    t2h, t2n = tl.t2_offset(t)
    assert np.array_equal(e2h, t2h) and e2 == t2n


def test_cr_tiled_oracle_equals_the_untiled_dense_evaluation(oracle):
    ref = None
    for ts, restricted in ((1, True), (2, True), (3, True), (8, True), (2, False)):
        st, cr, d = _inputs(ts, restricted)
        r = oracle.cr_ccsd_t(st, cr)
        dense = np.array(d.dense_reference()[:4])
        assert np.max(np.abs(r["sums"] - dense)) <= 1e-16, (ts, restricted, r["sums"], dense)
        if ref is None:
            ref = r
            assert abs(ref["sums"][0]) > 1e-6 and abs(ref["sums"][2]) > 1e-7 and abs(ref["sums"][1] - ref["sums"][0]) > 1e-6
            assert abs(cr.den0) > 1e-3
        assert np.max(np.abs(r["sums"] - ref["sums"])) <= 1e-16                 # tile-size / spin-adaptation invariance
        assert abs(r["e1"] - ref["e1"]) <= 1e-16 and abs(r["e2"] - ref["e2"]) <= 1e-16


def test_cr_moment_tile_is_antisymmetric_dense_tensor_and_reduces_to_the_t_doubles(oracle):
    """Tile by tile: the oracle's `moment 2,3` and `denominator` tiles are the corresponding blocks of the dense
    antisymmetric tensors; with amplitudes scaled by s the moment differs from the (T) doubles tile by O(s) relative."""
    from oracle import cr_dense
    t = tl.make_tiling(OCC, VIRT, 2)
    st, cr, d = _inputs(2)
    S, D, M, E = d.six_index()
    checked = 0
    for tup in oracle.task_list(t)[::7]:
        tup = [int(x) for x in tup[:6]]
        _, m, e = oracle.cr_tuple(st, cr, tup)
        s_ref, d_ref = oracle.tuple_tiles(st, tup)[:2]
        ix = np.ix_(d._pidx(tup[0]), d._pidx(tup[1]), d._pidx(tup[2]), d._hidx(tup[3]), d._hidx(tup[4]), d._hidx(tup[5]))
        assert np.max(np.abs(m - M[ix])) <= 1e-16 and np.max(np.abs(e - E[ix])) <= 1e-16
        assert np.max(np.abs(d_ref - D[ix])) <= 1e-16 and np.max(np.abs(s_ref - S[ix])) <= 1e-16
        checked += 1
    assert checked >= 5
    rel = []
    for sc in (1e-2, 1e-3):
        dd = cr_dense.Dense(t, t_scale=sc)
        _, D2, M2, _ = dd.six_index()
        rel.append(np.max(np.abs(M2 - D2)) / np.max(np.abs(D2)))
    assert rel[0] < 1e-2 and 8.0 < rel[0] / rel[1] < 12.0


# ------------------------------------------------------------------------------------------------------------------
# GPU: nwc_triples_set_cr / nwc_triples_run_cr.  Default: ONE dual-energy tuple per task through the LAMBDA instantiation
# of the fused kernel (M and D contracted once each; E formed in registers after M has been consumed).  NWC_CR_TWO_PASS=1:
# two two-sided tuples per task (numerators with M as the side-0 tile, denominators with the E outer products bound to it).
# ------------------------------------------------------------------------------------------------------------------
def _sorted_rows(tr, pt):
    order = sorted(range(len(pt)), key=lambda i: tuple(int(x) for x in tr.task_list()[i][:6]))
    return pt[order]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,ts,restricted,intorb", [(None, 2, True, False), (None, 3, True, True),
                                                         (None, 2, False, False), ("h2o_ccpvdz_c2v", 20, True, False)])
def test_cr_ccsd_t_gpu_matches_oracle(oracle, shape, ts, restricted, intorb):
    from nwchem_b200 import capi
    st, cr, d = _inputs(ts, restricted, shape)
    ref = oracle.cr_ccsd_t(st, cr)
    tr = capi.Triples(0)
    if intorb:
        tr.set_state_2eorb(st)
    else:
        tr.set_state(dataclasses.replace(st, orb=None))
    tr.set_cr(cr)
    sums, pt = tr.run_cr(per_task=True)
    got = _sorted_rows(tr, pt)
    e1, e2 = tr.cr_energies(sums, cr.den0)
    tr.close()
    scale = np.max(np.abs(ref["sums"]))
    assert np.max(np.abs(sums - ref["sums"])) <= 1e-12 * max(1.0, scale / 1e-3), (sums, ref["sums"])
    assert np.max(np.abs(sums - ref["sums"]) / np.abs(ref["sums"])) <= 1e-10
    assert np.max(np.abs(got - ref["per_task"])) <= 1e-13
    assert abs(e1 - ref["e1"]) <= 1e-12 and abs(e2 - ref["e2"]) <= 1e-12


@pytest.mark.gpu
def test_cr_block_partition_sums_to_total(oracle):
    from nwchem_b200 import capi
    st, cr, d = _inputs(20, True, "h2o_ccpvdz_c2v")
    ref = oracle.cr_ccsd_t(st, cr)
    tr = capi.Triples(0)
    tr.set_state(dataclasses.replace(st, orb=None))
    tr.set_cr(cr)
    sums, pt = tr.run_cr(per_task=True)
    parts = [tr.run_cr_partition(r, 3, per_task=True) for r in range(3)]
    again, _ = tr.run_cr(per_task=True)
    tr.close()
    assert np.max(np.abs(sums - ref["sums"])) <= 1e-12
    assert np.array_equal(sums, again)                                            # bitwise reproducible
    assert np.max(np.abs(sum(p[0] for p in parts) - sums)) <= 1e-14
    assert np.max(np.abs(sum(p[1] for p in parts) - pt)) <= 1e-15


@pytest.mark.gpu
def test_cr_ragged_and_random_blocks(oracle):
    """Stores without any permutational symmetry (every block iid) on a ragged tiling: the library must follow the
    reference block by block (which block, which element order, which sign), not merely the antisymmetric algebra."""
    from nwchem_b200 import capi
    t = tl.make_tiling([5], [11], 6)             # tiles 5 | 5,6 per spin: ragged against the 4-wide sub-tiles
    st = synth.random_blocks(t, seed=11)
    rng = np.random.default_rng(5)
    from oracle import cr_dense
    n1h, n1 = tl.cr_n1_offset(t); n2h, n2 = tl.cr_n2_offset(t); e2h, e2 = tl.cr_e2_offset(t)
    cr = cr_dense.CRStores(n1h, rng.uniform(-1, 1, n1) * 0.1, n2h, rng.uniform(-1, 1, n2) * 0.1, e2h,
                           rng.uniform(-1, 1, e2) * 0.02, 0.0)
    ref = oracle.cr_ccsd_t(st, cr)
    tr = capi.Triples(0)
    tr.set_state(st)
    tr.set_cr(cr)
    sums, pt = tr.run_cr(per_task=True)
    got = _sorted_rows(tr, pt)
    tr.close()
    assert np.max(np.abs(ref["per_task"])) > 1e-8
    assert np.max(np.abs(got - ref["per_task"])) <= 1e-12 * max(1.0, np.max(np.abs(ref["per_task"])))
    assert np.max(np.abs(sums - ref["sums"]) / np.abs(ref["sums"])) <= 1e-10


@pytest.mark.gpu
def test_cr_one_pass_equals_two_pass(oracle, monkeypatch):
    """The dual-tuple form against the two-pass form (same kernels as Lambda-CCSD(T), D contracted twice), per task, on the
    H2O table and on a ragged tiling with unsymmetric blocks; partition pieces of the dual form add up."""
    from nwchem_b200 import capi
    from oracle import cr_dense
    for which in ("h2o", "ragged"):
        if which == "h2o":
            st, cr, d = _inputs(20, True, "h2o_ccpvdz_c2v")
            st = dataclasses.replace(st, orb=None)
        else:
            t = tl.make_tiling([5], [11], 6)
            st = synth.random_blocks(t, seed=11)
            rng = np.random.default_rng(5)
            n1h, n1 = tl.cr_n1_offset(t); n2h, n2 = tl.cr_n2_offset(t); e2h, e2 = tl.cr_e2_offset(t)
            cr = cr_dense.CRStores(n1h, rng.uniform(-1, 1, n1) * 0.1, n2h, rng.uniform(-1, 1, n2) * 0.1, e2h,
                                   rng.uniform(-1, 1, e2) * 0.02, 0.0)
        tr = capi.Triples(0)
        tr.set_state(st)
        tr.set_cr(cr)
        monkeypatch.delenv("NWC_CR_TWO_PASS", raising=False)
        s1, p1 = tr.run_cr(per_task=True)
        parts = [tr.run_cr_partition(r, 2, per_task=True) for r in range(2)]
        monkeypatch.setenv("NWC_CR_TWO_PASS", "1")
        s2, p2 = tr.run_cr(per_task=True)
        monkeypatch.delenv("NWC_CR_TWO_PASS", raising=False)
        tr.close()
        scale = max(1.0, np.max(np.abs(p2)))
        assert np.max(np.abs(p1 - p2)) <= 1e-13 * scale, which
        assert np.max(np.abs(s1 - s2)) <= 1e-12 * max(1.0, np.max(np.abs(s2)))
        assert np.max(np.abs(sum(p[1] for p in parts) - p1)) <= 1e-13 * scale


# ------------------------------------------------------------------------------------------------------------------
# CR-EOMCCSD(T) (src/tce/cr-eomccsd_t/cr_eomccsd_t.F:325-493)
# ------------------------------------------------------------------------------------------------------------------
def test_creom_kernels_with_literal_factors_are_multiples_of_the_cr_kernels(oracle):
    """cre_t_K with the factor lists of creomsd_t_n2_mem_2 / _4 == -2 x / +1 x sd_t_d2cp_K (what the oracle's and the
    library's treatment of those routines as scaled cr_ccsd_t_N_2 rests on)."""
    rng = np.random.default_rng(1)
    dims = (3, 2, 4, 2, 3, 2)          # h3d,h2d,h1d,p6d,p5d,p4d
    kd = 5
    t2sub = rng.uniform(-1, 1, kd * dims[5] * dims[2] * dims[1])
    v2sub = rng.uniform(-1, 1, kd * dims[0] * dims[3] * dims[4])
    for k0 in range(9):
        a, b, d = oracle.cre_t_vs_d2cp(k0, dims, kd, t2sub, v2sub)
        assert np.max(np.abs(d)) > 0.1
        assert np.max(np.abs(a + 2.0 * d)) <= 1e-14 and np.max(np.abs(b - d)) <= 1e-14


def test_creom_tiled_oracle_equals_the_untiled_dense_evaluation(oracle):
    from oracle import cr_dense
    ref = {}
    for ts, restricted, r0 in ((1, True, 0.37), (2, True, 0.37), (3, True, 0.0), (3, True, 0.37), (2, False, 0.37), (2, False, 0.0)):
        t = tl.make_tiling(OCC, VIRT, ts, restricted)
        st = synth.physical(t)
        d = cr_dense.DenseEOM(t, r0=r0)
        cr, q = d.stores()
        r = oracle.cr_eomccsd_t(st, cr, q)
        dense = np.array(d.dense_reference())
        assert np.max(np.abs(r["sums"] - dense)) <= 1e-16, (ts, restricted, r0, r["sums"], dense)
        assert np.min(np.abs(r["sums"])) > 1e-7
        if r0 in ref:
            assert np.max(np.abs(r["sums"] - ref[r0])) <= 1e-16
        ref[r0] = r["sums"]
    assert np.max(np.abs(ref[0.0] - ref[0.37])) > 1e-6          # the r0 terms matter


@pytest.mark.gpu
@pytest.mark.parametrize("shape,ts,restricted,r0", [(None, 2, True, 0.37), (None, 3, True, 0.0), (None, 2, False, 0.37),
                                                     ("h2o_ccpvdz_c2v", 20, True, 0.37)])
def test_cr_eomccsd_t_gpu_matches_oracle(oracle, shape, ts, restricted, r0):
    from nwchem_b200 import capi
    from oracle import cr_dense
    t = tl.make_tiling(OCC, VIRT, ts, restricted) if shape is None else synth.shape_tiling(shape, tilesize=ts, restricted=restricted)
    st = synth.physical(t)
    cr, q = cr_dense.DenseEOM(t, r0=r0).stores()
    ref = oracle.cr_eomccsd_t(st, cr, q)
    tr = capi.Triples(0)
    tr.set_state(st)
    if abs(r0) >= 1e-7:
        tr.set_cr(cr)
    tr.set_creom(q)
    sums, pt = tr.run_creom(per_task=True)
    got = _sorted_rows(tr, pt)
    parts = [tr.run_creom_partition(r, 2, per_task=True) for r in range(2)]
    import os
    os.environ["NWC_CREOM_COMPOSED"] = "1"      # A/B: the form composed of a plain tuple and two unit-denominator tuples
    try:
        sums_c, pt_c = tr.run_creom(per_task=True)
    finally:
        del os.environ["NWC_CREOM_COMPOSED"]
    tr.close()
    scale = max(1.0, np.max(np.abs(ref["per_task"])))
    assert np.max(np.abs(pt_c - pt)) <= 1e-12 * scale
    assert np.max(np.abs(got - ref["per_task"])) <= 1e-12 * scale
    assert np.max(np.abs(sums - ref["sums"]) / np.abs(ref["sums"])) <= 1e-9
    assert np.max(np.abs(sum(p[0] for p in parts) - sums)) <= 1e-13
    assert np.max(np.abs(sum(p[1] for p in parts) - pt)) <= 1e-13 * scale


@pytest.mark.gpu
def test_cr_eomccsd_t_ragged_and_random_blocks(oracle):
    from nwchem_b200 import capi
    from oracle import cr_dense
    t = tl.make_tiling([5], [11], 6)
    st = synth.random_blocks(t, seed=11)
    rng = np.random.default_rng(5)
    n1h, n1 = tl.cr_n1_offset(t); n2h, n2 = tl.cr_n2_offset(t); e2h, e2 = tl.cr_e2_offset(t)
    r = lambda n, s: rng.uniform(-1, 1, n) * s
    cr = cr_dense.CRStores(n1h, r(n1, 0.1), n2h, r(n2, 0.1), e2h, r(e2, 0.02), 0.0)
    q = cr_dense.CREOMStores(st.t1_hash, r(len(st.t1), 0.05), st.t2_hash, r(len(st.t2), 0.02), n1h, r(n1, 0.1), n2h, r(n2, 0.1),
                             n1h, r(n1, 0.1), n2h, r(n2, 0.1), e2h, r(e2, 0.02), -0.6, 0.3)
    ref = oracle.cr_eomccsd_t(st, cr, q)
    tr = capi.Triples(0)
    tr.set_state(st)
    tr.set_cr(cr)
    tr.set_creom(q)
    sums, pt = tr.run_creom(per_task=True)
    got = _sorted_rows(tr, pt)
    tr.close()
    scale = max(1.0, np.max(np.abs(ref["per_task"])))
    assert np.max(np.abs(ref["per_task"])) > 1e-8
    assert np.max(np.abs(got - ref["per_task"])) <= 1e-11 * scale
    assert np.max(np.abs(sums - ref["sums"]) / np.abs(ref["sums"])) <= 1e-9


@pytest.mark.gpu
def test_cr_sharded_pphp_intermediate_two_contexts_one_gpu(oracle):
    """nwc_triples_set_cr_sharded: the pphp intermediate dealt block-wise over two "ranks" (two contexts on one GPU, peer
    pointers exchanged in-process; remote blocks pulled into the batch arena).  Each runs its piece of the block
    partition; the pieces add up to the replicated run bit for bit (a pulled block is a copy), which matches the oracle.
    CR-EOMCCSD(T) reads the same sharded store for its r0 term."""
    from nwchem_b200 import capi
    from oracle import cr_dense
    t = synth.shape_tiling("h2o_ccpvdz_c2v")
    st = synth.physical(t)
    cr, q = cr_dense.DenseEOM(t, r0=0.37).stores()
    ref = oracle.cr_ccsd_t(st, cr)
    ref_eom = oracle.cr_eomccsd_t(st, cr, q)
    one = capi.Triples(0)
    one.set_state(st)
    one.set_cr(cr)
    s1, p1 = one.run_cr(per_task=True)
    one.close()
    ctx = []
    for r in range(2):
        tr = capi.Triples(0)
        tr.set_state(st)
        tr.set_cr_sharded(dataclasses.replace(cr, n2=synth.shard_store(cr.n2_hash, cr.n2, r, 2)), r, 2)
        ctx.append(tr)
    ctx[0].cr_set_peer_ptr(1, ctx[1].cr_shard_ptr())
    ctx[1].cr_set_peer_ptr(0, ctx[0].cr_shard_ptr())
    parts = [ctx[r].run_cr_partition(r, 2, per_task=True) for r in range(2)]
    for tr in ctx:
        tr.set_creom(q)
    eom = [ctx[r].run_creom_partition(r, 2, per_task=True) for r in range(2)]
    for tr in ctx:
        tr.close()
    assert np.max(np.abs(s1 - ref["sums"])) <= 1e-12
    assert np.max(np.abs(parts[0][1] + parts[1][1] - p1)) <= 1e-15
    assert np.max(np.abs(parts[0][0] + parts[1][0] - ref["sums"])) <= 1e-12
    tot = eom[0][0] + eom[1][0]
    assert np.max(np.abs(tot - ref_eom["sums"]) / np.abs(ref_eom["sums"])) <= 1e-9


@pytest.mark.gpu
def test_sibling_corrections_reduce_to_the_golden_t_corrections_on_the_gpu():
    """The limits of tests/test_qa_h2o.py::test_sibling_restatements_reduce_to_the_golden_t_corrections through the CUDA
    library: Lambda-CCSD(T) with lambda := T^+, CR-CCSD(T) with undressed intermediates and CR-EOMCCSD(T) with r0 = 1,
    omega = 0 each contain the (T) correction, and on the first-principles amplitudes of the QA case that is the golden
    CCSD[T] / CCSD(T) of QA/tests/tce_ccsd_t_h2o/tce_ccsd_t_h2o.out:3374,3376."""
    from nwchem_b200 import capi
    from oracle import h2o_ccsd as h, cr_dense
    from test_qa_h2o import _lambda_from_t, _bare_cr_stores, TOL
    r = h.load()
    st = dataclasses.replace(h.qa_stores(r, tilesize=20, c2v=True, intorb=False), orb=None)
    g1, g2 = h.QA["t_bracket"], h.QA["t_paren"]
    lam = _lambda_from_t(st)
    cr = _bare_cr_stores(h, r, st)
    z = lambda a: np.zeros_like(a)
    q = cr_dense.CREOMStores(st.t1_hash, z(st.t1), st.t2_hash, z(st.t2), cr.n1_hash, z(cr.n1), cr.n2_hash, z(cr.n2),
                             cr.n1_hash, z(cr.n1), cr.n2_hash, z(cr.n2), cr.e2_hash, z(cr.e2), 1.0, 0.0)
    tr = capi.Triples(0)
    tr.set_state(st)
    tr.set_lambda(lam)
    le1, le2 = tr.run_lambda()[:2]
    tr.set_cr(cr)
    cs = tr.run_cr()
    cs = cs[0] if isinstance(cs, tuple) else cs
    tr.set_creom(q)
    es = tr.run_creom()
    es = es[0] if isinstance(es, tuple) else es
    tr.close()
    assert abs(le1 - g1) <= 2 * TOL and abs(le2 - g2) <= 2 * TOL, (le1, le2)
    assert abs(cs[0] - g1) <= 2 * TOL and abs(cs[1] - g2) <= 2 * TOL, cs
    assert abs(es[0] - g1) <= 2 * TOL, es


@pytest.mark.gpu
def test_cr_ccsd_t_gpu_on_the_glycine_qa_case(oracle):
    """CR-CCSD(T) through the CUDA library on the inputs of QA/tests/tce_lr_ccsd_t (glycine / STO-3G, ragged 7 + 8 hole
    tiles), whose moment and t1 (x) t2 tiles the oracle reproduces the reference's golden LR-CCSD(T) energies with
    (tests/test_qa_lr.py): per tuple and in total against that oracle."""
    from nwchem_b200 import capi
    from oracle import h2o_ccsd as h, cr_dense
    r = h.load(h.FIXTURE_GLYCINE)
    st = h.qa_stores(r, tilesize=10, c2v=False)
    cr = cr_dense.Dense(st.t, dense=(15, 10, r["t1s"], r["t2s"], r["eri_mo"])).stores()
    ref = oracle.cr_ccsd_t(st, cr)
    tr = capi.Triples(0)
    tr.set_state(st)
    tr.set_cr(cr)
    sums, pt = tr.run_cr(per_task=True)
    got = _sorted_rows(tr, pt)
    tr.close()
    assert np.max(np.abs(got - ref["per_task"])) <= 1e-12
    assert np.max(np.abs(sums - ref["sums"])) <= 1e-12
    e1, e2 = capi.Triples.cr_energies(sums, cr.den0)
    assert abs(e1 - ref["e1"]) <= 1e-12 and abs(e2 - ref["e2"]) <= 1e-12


@pytest.mark.gpu
def test_cr_ccsd_t_gpu_gives_the_published_h2o_dz_energy_at_2re(oracle):
    """The full CR-CCSD(T) energy through the CUDA library on the H2O / DZ full-CI benchmark at 2 R_e (tests/test_lit_h2o_dz.py):
    1.830 millihartree above full CI where CCSD(T) is 7.699 below -- the four sums and the scalar den0 all matter."""
    from nwchem_b200 import capi
    from oracle import h2o_ccsd as h, cr_dense
    r = h.generate_h2o_dz(2.0)
    st = h.qa_stores(r, tilesize=4, c2v=False)
    cr = cr_dense.Dense(st.t, dense=(5, 9, r["t1s"], r["t2s"], r["eri_mo"])).stores()
    lit = h.H2O_DZ_LIT[2.0]
    tr = capi.Triples(0)
    tr.set_state(st)
    e_t = tr.run()
    tr.set_cr(cr)
    sums = tr.run_cr()
    tr.close()
    e2 = capi.Triples.cr_energies(sums, cr.den0)[1]
    ccsd = float(r["escf"]) + float(r["ecc"])
    assert abs(ccsd + e2 - (lit["fci"] + 1e-3 * lit["cr_ccsd_t"])) <= 1.5e-6
    assert abs(ccsd + e_t[1] - (lit["fci"] + 1e-3 * lit["ccsd_t"])) <= 1.5e-6
    ref = oracle.cr_ccsd_t(st, cr)
    assert np.max(np.abs(sums - ref["sums"])) <= 1e-12
