"""CPU tests that pin the oracle (oracle/triples_oracle.c) -- no GPU needed.

Pins available without a Fortran compiler: the QA tile table, tile-size invariance of
E[T]/E(T) on antisymmetric synthetic amplitudes, an independent second formulation of the doubles
(sort -> GEMM -> sortacc), and the survey's dispatch/flop counts.
"""
import ctypes as C
import json
import os
import numpy as np
import pytest
from nwchem_b200 import synth, tiling as tl

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_h2o_tile_table_matches_qa_output(oracle):
    # QA/tests/tce_ccsd_t_h2o/tce_ccsd_t_h2o.out:644-659
    g = json.load(open(os.path.join(GOLD, "h2o_tile_table.json")))
    t = synth.shape_tiling("h2o_ccpvdz_c2v")
    assert t.range.tolist() == g["size"]
    assert t.offset.tolist() == g["offset"]
    assert t.alpha.tolist() == g["alpha"]
    assert [int(s) for s in t.spin] == g["spin"]
    assert [int(s) for s in t.sym] == g["irrep"]
    # the oracle's own restatement of tce_tile.F:330-357 agrees with the host-side one
    for n in range(0, 60):
        for isize in (1, 3, 8, 20, 24, 40):
            assert oracle.tile_group(n, isize) == tl.tile_group(n, isize)
    assert oracle.tile_group(870, 40) == tl.tile_group(870, 40) and len(tl.tile_group(870, 40)) == 22


@pytest.mark.parametrize("shape", ["h2o_ccpvdz_c2v", "h2o_ccpvdz_c1"])
def test_tile_size_invariance(oracle, shape):
    ref = None
    for ts in (3, 5, 8, 20):
        st = synth.physical(synth.shape_tiling(shape, tilesize=ts))
        r = oracle.ccsd_t(st)
        if ref is None:
            ref = r
        assert abs(r["e1"] - ref["e1"]) <= 1e-12 * abs(ref["e1"])
        assert abs(r["e2"] - ref["e2"]) <= 1e-12 * abs(ref["e2"])
    assert ref["e1"] < 0 and ref["e2"] < 0


def test_unrestricted_equals_restricted(oracle):
    """The same closed-shell tensors treated as UHF (beta tiles explicit, factor 1, no k_alpha mapping) must give the
    RHF-restricted energies (factor 2, spin-sum<=8 task filter, tce_restricted_2/4): ccsd_t_dot.F:52-56."""
    for shape, ts in (("h2o_ccpvdz_c2v", 20), ("h2o_ccpvdz_c1", 7)):
        r = oracle.ccsd_t(synth.physical(synth.shape_tiling(shape, tilesize=ts)))
        u = oracle.ccsd_t(synth.physical(synth.shape_tiling(shape, tilesize=ts, restricted=False)))
        assert len(u["tasks"]) == 2 * len(r["tasks"])
        assert abs(u["e1"] - r["e1"]) <= 1e-12 * abs(r["e1"]) and abs(u["e2"] - r["e2"]) <= 1e-12 * abs(r["e2"])


def test_restartable_t_table_and_resume(oracle):
    """ccsd_t_restart.F: the per-outer-virtual-tile table sums to E(T) of the plain driver; a run interrupted after
    three outer tiles and resumed from the saved (begin, table) reproduces the uninterrupted table exactly."""
    st = synth.physical(synth.shape_tiling("h2o_ccpvdz_c2v"))
    full = oracle.ccsd_t(st)
    begin, table, t_energy, done = oracle.ccsd_t_restart(st)
    assert begin == st.t.nvab + 1 and done == st.t.nvab
    assert abs(t_energy - full["e2"]) <= 1e-13 * abs(full["e2"])
    # each entry is the sum of the per-task energies of the tuples whose first virtual tile is that outer tile
    by_outer = np.zeros(st.t.nvab)
    for task, e in zip(full["tasks"], full["per_task"]):
        by_outer[int(task[0]) - st.t.noab - 1] += e[1]
    assert np.allclose(table, by_outer, rtol=0, atol=1e-15)
    b1, t1, _, d1 = oracle.ccsd_t_restart(st, max_outer=3)
    assert (b1, d1) == (4, 3) and np.all(t1[3:] == 0.0)
    b2, t2, te2, d2 = oracle.ccsd_t_restart(st, begin=b1, table=t1)
    assert b2 == st.t.nvab + 1 and d2 == st.t.nvab - 3
    # (the oracle's OpenMP reduction in ccsd_t_dot is not bitwise reproducible run to run, hence 1e-15, not ==)
    assert np.allclose(t2, table, rtol=0, atol=1e-15) and abs(te2 - t_energy) <= 1e-15
    # nothing left to do: the table is returned untouched
    b3, t3, te3, d3 = oracle.ccsd_t_restart(st, begin=b2, table=t2)
    assert d3 == 0 and np.array_equal(t3, t2) and te3 == te2


@pytest.mark.parametrize("shape,ts,restricted", [("h2o_ccpvdz_c2v", 20, True), ("h2o_ccpvdz_c2v", 20, False),
                                                   ("h2o_ccpvdz_c1", 7, True)])
def test_2eorb_storage_reproduces_spin_orbital_v2(oracle, shape, ts, restricted):
    """`2eorb` (intorb) path, get_block_ind.F:818-1538 + tce_hash.F:1-135 restated: the same spatial integrals stored
    spin-free over the alpha tiles (tce_mo2e_offset_intorb.F layout, checkpointed table with idiv2e = 2) must give
    every spin-orbital V2 block the (T) path reads, bit for bit, and therefore the same E[T] / E(T)."""
    import dataclasses
    t = synth.shape_tiling(shape, tilesize=ts, restricted=restricted)
    st = synth.physical(t, intorb=True)
    blocks, size = tl.v2orb_blocks(st.orb.a)
    assert size == len(st.orb.v2orb) and int(st.orb.v2orb_hash[0]) in (2, 3)
    # tce_hash_v2 finds every stored block at the offset the builder gave it; an absent key is reported
    assert all(oracle.hash_v2(st.orb, b[4]) == b[5] for b in blocks)
    assert oracle.hash_v2(st.orb, 10 ** 9) == -1
    spins = set()
    for key, off in synth._iter_hash(st.v2_hash):
        g3b, g4b, g1b, g2b = tl.decode_v2_key(t, key)
        n = t.r(g3b) * t.r(g4b) * t.r(g1b) * t.r(g2b)
        assert np.array_equal(oracle.v2_block_intorb(st, g3b, g4b, g1b, g2b), st.v2[off:off + n])
        spins.add(tuple(int(t.spin[b - 1]) for b in (g3b, g4b, g1b, g2b)))
    assert (1, 1, 1, 1) in spins and (1, 2, 1, 2) in spins
    if not restricted:
        assert {(2, 2, 2, 2), (1, 2, 2, 1), (2, 1, 1, 2)} <= spins
    a = oracle.ccsd_t(dataclasses.replace(st, orb=None))
    b = oracle.ccsd_t(st)
    # identical blocks -> identical tiles; only the OpenMP reduction order of ccsd_t_dot may differ
    assert abs(a["e1"] - b["e1"]) <= 1e-14 * abs(a["e1"]) and abs(a["e2"] - b["e2"]) <= 1e-14 * abs(a["e2"])


def test_golden_energies(oracle):
    # generated by tests/golden/make_golden.py with this oracle (regression pin of the restatement)
    g = json.load(open(os.path.join(GOLD, "oracle_energies.json")))
    for case in g["cases"]:
        st = synth.physical(synth.shape_tiling(case["shape"], tilesize=case["tilesize"]))
        r = oracle.ccsd_t(st)
        assert len(r["tasks"]) == case["tasks"]
        assert abs(r["e1"] - case["e1"]) <= 1e-12 * abs(case["e1"])
        assert abs(r["e2"] - case["e2"]) <= 1e-12 * abs(case["e2"])
        assert r["counts"].flops == case["flops"]


def test_golden_restart_table_and_2eorb_offsets(oracle):
    # regression pins generated by tests/golden/make_golden.py (restart table: ccsd_t_restart.F; k_b2am and the
    # checkpointed k_v2_alpha_offset table: tce_tile.F:1156-1212, tce_mo2e_offset_intorb.F)
    g = json.load(open(os.path.join(GOLD, "h2o_restart_2eorb.json")))
    st = synth.physical(synth.shape_tiling(g["shape"], tilesize=g["tilesize"]), intorb=True)
    _, table, t_energy, _ = oracle.ccsd_t_restart(st)
    assert np.allclose(table, g["restart_table"], rtol=1e-12, atol=1e-16)
    assert abs(t_energy - g["restart_t_energy"]) <= 1e-12 * abs(t_energy)
    a = st.orb.a
    assert [int(x) for x in a.b2am] == g["b2am"] and [int(x) for x in a.range_alpha] == g["range_alpha"]
    assert [int(x) for x in a.sym_alpha] == g["sym_alpha"]
    assert [int(x) for x in st.orb.v2orb_hash] == g["v2orb_hash"] and len(st.orb.v2orb) == g["v2orb_size"]


def test_survey_flop_and_call_counts(oracle):
    # dry run of the dispatch logic, no arithmetic
    class D:
        pass
    expect = {"microbench_t40": (2, 1.126e13, None), "uracil_augccpvdz": (110, 2.01e14, None),
              "h2o10_augccpvtz": (7590, 4.52e17, (46046, 62744, 1380368))}
    for name, (ntask, flops, calls) in expect.items():
        t = synth.shape_tiling(name)
        d = D(); d.t = t
        d.t1_hash = d.t2_hash = d.v2_hash = np.zeros(3, np.int64); d.t1 = d.t2 = d.v2 = np.zeros(1)
        c, keep = oracle.make_ctx(d)
        tasks = oracle.task_list(t)
        assert len(tasks) == ntask
        f, cs = 0.0, [0, 0, 0]
        for tup in tasks:
            cnt = oracle.count_tuple(c, tup[:6], keep)
            f += cnt.flops; cs[0] += cnt.calls_s1; cs[1] += cnt.calls_d1; cs[2] += cnt.calls_d2
        assert abs(f - flops) / flops < 5e-3
        if calls:
            assert tuple(cs) == calls


def test_task_list_sorted_heaviest_first(oracle):
    t = synth.shape_tiling("h2o_ccpvdz_c2v", tilesize=5)
    kl = oracle.task_list(t)
    w = kl[:, 6]
    # 16-band bucket sort (ccsd_t_neword.F:168-174): band index never increases
    wl_min, wl_max = w.min(), w.max()
    band = np.array([max(ii for ii in range(0, 17) if (ii == 0 or x > wl_min + ((wl_max - wl_min) * (ii - 1)) // 16)) for x in w])
    assert np.all(np.diff(band) <= 0)
    assert len({tuple(r[:6]) for r in kl}) == len(kl)


def test_second_formulation_doubles(oracle):
    """One (row,p7b) pair of the Sum(p7) family through sort->GEMM->sortacc_6 (ccsd_t_doubles.F:195-267)
    equals kernel sd_t_d2_1 (ccsd_t_kernels_omp.F:857-893); same for d1_1."""
    l = oracle.lib()
    rng = np.random.default_rng(7)
    h3d, h2d, h1d, p6d, p5d, p4d, kd = 3, 4, 2, 5, 3, 4, 6
    dims = (h3d, h2d, h1d, p6d, p5d, p4d)
    n = int(np.prod(dims))
    PD = C.POINTER(C.c_double); L = C.c_long
    pd = lambda a: a.ctypes.data_as(PD)
    # ---- d2_1: t2sub(p7,p4,h1,h2), v2sub(p7,h3,p6,p5)
    t2sub = rng.standard_normal(kd * p4d * h1d * h2d); v2sub = rng.standard_normal(kd * h3d * p6d * p5d)
    t3 = np.zeros(n); oracle.kernel(2, 1, dims, kd, t3, t2sub, v2sub)
    cs = np.zeros(p4d * h1d * h2d * h3d * p6d * p5d)
    l.ora_dgemm_tn(L(p4d * h1d * h2d), L(h3d * p6d * p5d), L(kd), pd(t2sub), pd(v2sub), pd(cs))
    # c_sort col-major (p4,h1,h2 | h3,p6,p5) == row-major-last-fastest dims (p5,p6,h3,h2,h1,p4);
    # target T3(h3,h2,h1,p6,p5,p4) == last-fastest dims (p4,p5,p6,h1,h2,h3)
    t3b = np.zeros(n)
    l.ora_tce_sortacc_6(pd(cs), pd(t3b), L(p5d), L(p6d), L(h3d), L(h2d), L(h1d), L(p4d), L(6), L(1), L(2), L(5), L(4), L(3), C.c_double(-1.0))
    assert np.allclose(t3, t3b, rtol=1e-13, atol=1e-13)
    # and against plain numpy
    A = t2sub.reshape(h2d, h1d, p4d, kd); B = v2sub.reshape(p5d, p6d, h3d, kd)
    ref = -np.einsum("jiak,cbhk->acbijh", A, B)  # [p4,p5,p6,h1,h2,h3]
    assert np.allclose(t3.reshape(p4d, p5d, p6d, h1d, h2d, h3d), ref, rtol=1e-13, atol=1e-13)
    # ---- d1_1: t2sub(h7,p4,p5,h1), v2sub(h3,h2,p6,h7)
    t2sub = rng.standard_normal(kd * p4d * p5d * h1d); v2sub = rng.standard_normal(h3d * h2d * p6d * kd)
    t3 = np.zeros(n); oracle.kernel(1, 1, dims, kd, t3, t2sub, v2sub)
    A = t2sub.reshape(h1d, p5d, p4d, kd); B = v2sub.reshape(kd, p6d, h2d, h3d)
    ref = -np.einsum("ibak,kcjh->abcijh", A, B)
    assert np.allclose(t3.reshape(p4d, p5d, p6d, h1d, h2d, h3d), ref, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("restricted", [True, False])
def test_second_formulation_singles_every_tuple(oracle, restricted):
    """The singles tile of EVERY tuple of the H2O C2v table through the original TCE-generated form
    (ccsd_t_singles.F:140-246: TCE_SORT_4(4,3,2,1), outer product, one TCE_SORTACC_6 per dispatch test with the
    permutation and sign written there) equals the nine loop kernels sd_t_s1_1..9 -- the layouts and signs of the
    kernels derived a second, independent way, including diagonal tuples where several tests fire on one row."""
    st = synth.physical(synth.shape_tiling("h2o_ccpvdz_c2v", restricted=restricted))
    tasks = oracle.task_list(st.t)
    fired = 0
    for row in tasks:
        tup = [int(x) for x in row[:6]]
        a = oracle.tuple_tiles(st, tup)[0]
        b = oracle.singles_tce(st, tup)
        assert np.allclose(a, b, rtol=0, atol=1e-16), tup
        fired += int(np.any(a != 0.0))
    assert fired > len(tasks) // 4


@pytest.mark.parametrize("restricted", [True, False])
def test_second_formulation_doubles_every_tuple(oracle, restricted):
    """The doubles tile of EVERY tuple of the H2O C2v table through ccsd_t_doubles.F's own formulation -- V2 sorted
    with the contracted index fastest (TCE_SORT_4 4,3,2,1 / 3,2,1,4), DGEMM('T','N'), and the eighteen TCE_SORTACC_6
    permutations and signs written in ccsd_t_doubles.F:206-267 and :457-520 -- equals the eighteen loop kernels
    sd_t_d1_1..9 / sd_t_d2_1..9 driven by offl_ccsd_t_doubles_l.F."""
    st = synth.physical(synth.shape_tiling("h2o_ccpvdz_c2v", restricted=restricted))
    tasks = oracle.task_list(st.t)
    scale = 0.0
    for row in tasks:
        tup = [int(x) for x in row[:6]]
        a = oracle.tuple_tiles(st, tup)[1]
        b = oracle.doubles_tce(st, tup)
        assert np.allclose(a, b, rtol=0, atol=1e-15 * max(1.0, float(np.abs(a).max()))), tup
        scale = max(scale, float(np.abs(a).max()))
    assert scale > 0.0


def test_all_27_kernels_against_numpy(oracle):
    """Each kernel == its declaration: triplesx(<declared order>) +-= tsub * v2sub (ccsd_t_kernels_omp.F)."""
    from nwchem_b200.kernel_tables import DECL, SIGN  # the product's own table, checked here against the oracle
    rng = np.random.default_rng(11)
    d = dict(h3=2, h2=3, h1=4, p6=3, p5=2, p4=5)
    kd = 4
    letters = dict(h3="a", h2="b", h1="c", p6="d", p5="e", p4="f")
    for fam in (0, 1, 2):
        for k in range(1, 10):
            decl = DECL[fam][k - 1]  # names fastest-first
            n = int(np.prod([d[x] for x in decl]))
            if fam == 0:
                ts = rng.standard_normal(d["p4"] * d["h1"]); vs = rng.standard_normal(d["h3"] * d["h2"] * d["p6"] * d["p5"])
                T = ts.reshape(d["h1"], d["p4"]); V = vs.reshape(d["p5"], d["p6"], d["h2"], d["h3"])
                full = np.einsum("cf,edba->abcdef", T, V)
            elif fam == 1:
                ts = rng.standard_normal(kd * d["p4"] * d["p5"] * d["h1"]); vs = rng.standard_normal(d["h3"] * d["h2"] * d["p6"] * kd)
                T = ts.reshape(d["h1"], d["p5"], d["p4"], kd); V = vs.reshape(kd, d["p6"], d["h2"], d["h3"])
                full = np.einsum("cefk,kdba->abcdef", T, V)
            else:
                ts = rng.standard_normal(kd * d["p4"] * d["h1"] * d["h2"]); vs = rng.standard_normal(kd * d["h3"] * d["p6"] * d["p5"])
                T = ts.reshape(d["h2"], d["h1"], d["p4"], kd); V = vs.reshape(d["p5"], d["p6"], d["h3"], kd)
                full = np.einsum("bcfk,edak->abcdef", T, V)
            t3 = np.zeros(n)
            oracle.kernel(fam, k, (d["h3"], d["h2"], d["h1"], d["p6"], d["p5"], d["p4"]), kd, t3, ts, vs)
            # declared order fastest-first -> numpy shape slowest-first
            got = t3.reshape([d[x] for x in reversed(decl)])
            sub = "".join(letters[x] for x in reversed(decl))
            exp = SIGN[fam][k - 1] * np.einsum("abcdef->" + sub, full)
            assert np.allclose(got, exp, rtol=1e-13, atol=1e-13), (fam, k)
