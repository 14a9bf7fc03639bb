"""The native tier's host driver against the oracle WITHOUT a GPU.

A trace context (nwc_triples_create_trace) runs the library's own driver logic -- csrc/host_driver.h walkers + the
NativeSink of csrc/native_abi.cu: which stored block a tuple reads, through which element strides, into which of the
reference's kernels, with which sign -- and records the operand descriptors it would hand to the engine.  This file
evaluates those records with numpy (each record is one call of a reference kernel: its declared index order and sign
come from nwchem_b200/kernel_tables.py) and compares the resulting t3 tiles with the oracle's, for (T), Lambda-CCSD(T)
and both passes of CR-CCSD(T).  What is left to the GPU tests is the engine (panels, fused kernel, reduction)."""
import ctypes as C
import dataclasses
import numpy as np
import pytest
from nwchem_b200 import capi, synth, tiling as tl
from nwchem_b200.kernel_tables import DECL, SIGN

OCC, VIRT = [2, 1], [3, 2]
NAMES = ("h1", "h2", "h3", "p4", "p5", "p6")          # permuted names in the order of nwc_trace_rec.sa/sb (tables.h N_*)


def _gather(ptr, offs):
    """doubles at address ptr + 8*offs (offs: integer array); reads the caller's / the library's host memory in place"""
    lo, hi = int(offs.min()), int(offs.max())
    span = np.ctypeslib.as_array((C.c_double * (hi - lo + 1)).from_address(int(ptr) + 8 * lo))
    return span[offs - lo]


def evaluate(recs):
    """Records of ONE tuple -> (side-0 doubles tile, side-1 tile, singles tile, factor, two_sided), tiles indexed
    [p4,p5,p6,h1,h2,h3] (C order == the physical T3(h3,h2,h1,p6,p5,p4)).  For a dual-energy tuple a sixth value follows:
    the tile of the side-0 outer products, which the kernel keeps apart from the side-0 contractions."""
    end = recs[-1]
    assert end.kind == 9 and all(r.kind != 9 for r in recs[:-1])
    dual = int(end.K) == 2
    R = [int(end.sa[q]) for q in range(6)]                                  # by physical position, h3 first
    shape = R[::-1]                                                         # numpy axes: p4,p5,p6,h1,h2,h3
    idx = [np.arange(R[q]).reshape([-1 if ax == 5 - q else 1 for ax in range(6)]) for q in range(6)]   # idx[q]: position q
    tiles = [np.zeros(shape), np.zeros(shape), np.zeros(shape), np.zeros(shape)]   # side 0, side 1, singles, dual: fourth tile
    for r in recs[:-1]:
        if r.kind == 3:                                                     # strides per physical position
            oa = sum(idx[q] * int(r.sa[q]) for q in range(6))
            ob = sum(idx[q] * int(r.sb[q]) for q in range(6))
            val = _gather(r.a, np.broadcast_to(oa, shape)) * _gather(r.b, np.broadcast_to(ob, shape))
            tiles[{0: 2, 1: 1, 2: 3 if dual else 0}[int(r.side)]] += -val if r.neg else val
            continue
        fam, k0 = int(r.kind), int(r.k0)
        decl = DECL[fam][k0]                                                # permuted name at each physical position
        name_of = [NAMES.index(decl[q]) for q in range(6)]
        oa = sum(idx[q] * int(r.sa[name_of[q]]) for q in range(6))
        ob = sum(idx[q] * int(r.sb[name_of[q]]) for q in range(6))
        oa = np.broadcast_to(oa, shape); ob = np.broadcast_to(ob, shape)
        if fam == 0:
            tiles[2] += SIGN[0][k0] * _gather(r.a, oa) * _gather(r.b, ob)
        else:
            acc = np.zeros(shape)
            for k in range(int(r.K)):
                acc += _gather(r.a, oa + k * int(r.ka)) * _gather(r.b, ob + k * int(r.kb))
            tiles[int(r.side)] += SIGN[fam][k0] * r.scale * acc
    if dual:
        return tiles[0], tiles[1], tiles[2], float(end.scale), True, tiles[3]
    return tiles[0], tiles[1], tiles[2], float(end.scale), bool(end.K)


def _eps_of(end, shape):
    """the six orbital-energy vectors the tuple was emitted with (end record), in the order p4,p5,p6,h1,h2,h3"""
    ptrs = [int(end.a), int(end.b), int(end.sb[2]), int(end.sb[3]), int(end.sb[4]), int(end.sb[5])]   # h1,h2,h3,p4,p5,p6
    rng = {0: shape[3], 1: shape[4], 2: shape[5], 3: shape[0], 4: shape[1], 5: shape[2]}
    v = [np.ctypeslib.as_array((C.c_double * rng[i]).from_address(ptrs[i])).copy() for i in range(6)]
    return [v[3], v[4], v[5], v[0], v[1], v[2]]


def _energies(t, tup, w_tile, d_tile, s_tile, factor, eps=None):
    """(sum f W D / Delta, sum f W (D + S) / Delta) as the fused kernel's energy pass forms them"""
    e = eps if eps is not None else [t.evl_sorted[t.offset[b - 1]:t.offset[b - 1] + t.range[b - 1]] for b in tup]
    delta = (-e[0][:, None, None, None, None, None] - e[1][None, :, None, None, None, None] - e[2][None, None, :, None, None, None]
             + e[3][None, None, None, :, None, None] + e[4][None, None, None, None, :, None] + e[5][None, None, None, None, None, :])
    return factor * np.sum(w_tile * d_tile / delta), factor * np.sum(w_tile * (d_tile + s_tile) / delta)


@pytest.mark.parametrize("ts,restricted", [(2, True), (3, True), (2, False)])
def test_t_driver_trace_matches_oracle_tiles(oracle, ts, restricted):
    t = tl.make_tiling(OCC, VIRT, ts, restricted)
    st = synth.physical(t)
    tr = capi.Triples(trace=True)
    tr.set_state(st)
    n = 0
    for tup in oracle.task_list(t)[::3]:
        tup = [int(x) for x in tup[:6]]
        recs, keep = tr.trace_tuple(tup, 0)
        d, _, s, f, two = evaluate(recs)
        s_ref, d_ref, e1, e2, _ = oracle.tuple_tiles(st, tup)
        assert not two
        assert np.max(np.abs(d - d_ref)) <= 1e-15 and np.max(np.abs(s - s_ref)) <= 1e-15
        g1, g2 = _energies(t, tup, d, d, s, f)
        assert abs(g1 - e1) <= 1e-15 and abs(g2 - e2) <= 1e-15
        n += 1
    tr.close()
    assert n >= 5


def test_t_driver_trace_random_blocks_ragged(oracle):
    """iid blocks (no permutational symmetry): the driver must pick the reference's block, element order and sign"""
    t = tl.make_tiling([5], [7], 4)
    st = synth.random_blocks(t, seed=3)
    tr = capi.Triples(trace=True)
    tr.set_state(st)
    for tup in oracle.task_list(t)[::4]:
        tup = [int(x) for x in tup[:6]]
        recs, keep = tr.trace_tuple(tup, 0)
        d, _, s, f, _ = evaluate(recs)
        s_ref, d_ref = oracle.tuple_tiles(st, tup)[:2]
        assert np.max(np.abs(d - d_ref)) <= 1e-13 and np.max(np.abs(s - s_ref)) <= 1e-14
    tr.close()


def test_lambda_driver_trace_matches_oracle_tiles(oracle):
    t = tl.make_tiling(OCC, VIRT, 2)
    st = synth.physical(t, intorb=True)      # the oracle's left-hand side reads <hh||pp>-type keys: it takes V2 from the 2eorb store
    lam = synth.physical_lambda(t)
    tr = capi.Triples(trace=True)
    tr.set_state(dataclasses.replace(st, orb=None))
    tr.set_lambda(lam)
    n = 0
    for tup in oracle.task_list(t)[::3]:
        tup = [int(x) for x in tup[:6]]
        recs, keep = tr.trace_tuple(tup, 1)
        td, yd, ys, f, two = evaluate(recs)
        e1, e2, td_ref, ys_ref, yd_ref = oracle.lambda_tuple(st, lam, tup, sorted=True)
        assert two
        assert np.max(np.abs(td - td_ref)) <= 1e-15
        assert np.max(np.abs(yd - yd_ref.transpose(3, 4, 5, 0, 1, 2))) <= 1e-15      # oracle: L3 order [h1,h2,h3,p4,p5,p6]
        assert np.max(np.abs(ys - ys_ref.transpose(3, 4, 5, 0, 1, 2))) <= 1e-15
        g1, g2 = _energies(t, tup, td, yd, ys, f)
        assert abs(g1 - e1) <= 1e-15 and abs(g2 - e2) <= 1e-15
        n += 1
    tr.close()
    assert n >= 5


@pytest.mark.parametrize("ts,restricted,kind", [(2, True, "physical"), (3, True, "physical"), (2, False, "physical"),
                                                (4, True, "random")])
def test_cr_driver_trace_matches_oracle_tiles(oracle, ts, restricted, kind):
    """Both CR-CCSD(T) passes: pass 0 = (M | D, S), pass 1 = (E | D, S); the four sums of cr_ccsd_t.F:176-207 follow."""
    from oracle import cr_dense
    if kind == "physical":
        t = tl.make_tiling(OCC, VIRT, ts, restricted)
        st = synth.physical(t)
        cr = cr_dense.Dense(t).stores()
    else:
        t = tl.make_tiling([5], [7], ts, restricted)
        st = synth.random_blocks(t, seed=3)
        rng = np.random.default_rng(9)
        n1h, n1 = tl.cr_n1_offset(t); n2h, n2 = tl.cr_n2_offset(t); e2h, e2 = tl.cr_e2_offset(t)
        cr = cr_dense.CRStores(n1h, rng.uniform(-1, 1, n1) * 0.1, n2h, rng.uniform(-1, 1, n2) * 0.1, e2h,
                               rng.uniform(-1, 1, e2) * 0.02, 0.0)
    tr = capi.Triples(trace=True)
    tr.set_state(st)
    tr.set_cr(cr)
    n = 0
    tol = 1e-15 if kind == "physical" else 1e-13
    for tup in oracle.task_list(t)[::3]:
        tup = [int(x) for x in tup[:6]]
        sums_ref, m_ref, e_ref = oracle.cr_tuple(st, cr, tup)
        s_ref, d_ref = oracle.tuple_tiles(st, tup)[:2]
        recs, keep = tr.trace_tuple(tup, 2)
        m, d, s, f, two = evaluate(recs)
        assert two and np.max(np.abs(m - m_ref)) <= tol and np.max(np.abs(d - d_ref)) <= tol and np.max(np.abs(s - s_ref)) <= tol
        num = _energies(t, tup, m, d, s, f)
        recs, keep = tr.trace_tuple(tup, 3)
        e, d2, s2, f2, two = evaluate(recs)
        assert two and np.max(np.abs(e - e_ref)) <= tol and np.array_equal(d2, d) and np.array_equal(s2, s) and f2 == f
        assert sum(1 for r in recs if r.kind == 3 and r.side == 2) <= 18 and len([r for r in recs if r.kind in (0, 3)]) <= 27
        den = _energies(t, tup, e, d, s, f)
        got = np.array([num[0], num[1], den[0], den[1]])
        assert np.max(np.abs(got - sums_ref)) <= 1e-13 * max(1.0, np.max(np.abs(sums_ref)))
        # the one-pass form (the default of nwc_triples_run_cr): one dual tuple holding all four tiles
        recs, keep = tr.trace_tuple(tup, 4)
        m1, d1, s1, f1, two, e1 = evaluate(recs)
        assert np.array_equal(m1, m) and np.array_equal(d1, d) and np.array_equal(s1, s) and np.array_equal(e1, e) and f1 == f
        n += 1
    tr.close()
    assert n >= 5


def test_trace_context_cannot_compute():
    t = tl.make_tiling(OCC, VIRT, 2)
    st = synth.physical(t)
    tr = capi.Triples(trace=True)
    tr.set_state(st)
    with pytest.raises(RuntimeError):
        tr.run()
    with pytest.raises(RuntimeError):
        tr.set_state_2eorb(synth.physical(t, intorb=True))
    tr.close()


@pytest.mark.parametrize("ts,restricted,r0", [(2, True, 0.37), (3, True, 0.0), (2, False, 0.37)])
def test_creom_driver_trace_matches_oracle(oracle, ts, restricted, r0):
    """CR-EOMCCSD(T): the three tuples the library emits per task (a plain tuple with shifted denominators, two two-sided
    tuples with unit denominators), evaluated from the trace with the energy formulas of the kernel, give the four sums of
    cr_eomccsd_t.F:455-464; the right / left tiles equal the oracle's."""
    from oracle import cr_dense
    t = tl.make_tiling(OCC, VIRT, ts, restricted)
    st = synth.physical(t)
    cr, q = cr_dense.DenseEOM(t, r0=r0).stores()
    tr = capi.Triples(trace=True)
    tr.set_state(st)
    if abs(r0) >= 1e-7:
        tr.set_cr(cr)
    tr.set_creom(q)
    n = 0
    for tup in oracle.task_list(t)[::3]:
        tup = [int(x) for x in tup[:6]]
        sums_ref, r_ref, l_ref = oracle.cr_eom_tuple(st, cr, q, tup)
        recs, keep = tr.trace_tuple(tup, 5)                         # X: plain, doubles = R, singles = L, denex
        rt, _, lt, f, two = evaluate(recs)
        assert not two and np.max(np.abs(rt - r_ref)) <= 1e-15 and np.max(np.abs(lt - l_ref)) <= 1e-15
        eps = _eps_of(recs[-1], rt.shape)
        ref_eps = [t.evl_sorted[t.offset[b - 1]:t.offset[b - 1] + t.range[b - 1]] for b in tup]
        assert np.allclose(eps[3], ref_eps[3] + q.excit, rtol=0, atol=1e-15) and all(np.array_equal(eps[i], ref_eps[i]) for i in (0, 1, 2, 4, 5))
        a, apc = _energies(t, tup, rt, rt, lt, f, eps)
        recs, keep = tr.trace_tuple(tup, 6)                         # Y1: side 0 = L, side 1 = R, singles = La, unit denominators
        l0, r1, la, f1, two = evaluate(recs)
        eps1 = _eps_of(recs[-1], rt.shape)
        assert two and np.array_equal(l0, lt) and np.array_equal(r1, rt) and f1 == f
        assert all(np.all(eps1[i] == (1.0 if i == 3 else 0.0)) for i in range(6))
        b, bpla = _energies(t, tup, l0, r1, la, f, eps1)
        recs, keep = tr.trace_tuple(tup, 7)                         # Y2: side 0 = L, no contraction, singles = Lb
        l2, z2, lb, f2, two = evaluate(recs)
        assert two and np.array_equal(l2, lt) and not np.any(z2) and np.max(np.abs(la + lb - lt)) <= 1e-16
        _, llb = _energies(t, tup, l2, z2, lb, f, _eps_of(recs[-1], rt.shape))
        got = np.array([a, b, apc - a, (bpla - b) + llb])
        assert np.max(np.abs(got - sums_ref)) <= 1e-13 * max(1.0, np.max(np.abs(sums_ref))), (got, sums_ref)
        # the one-tuple form (default of nwc_triples_run_creom): a dual-energy tuple, side 1 = R, singles = L, denex;
        # the kernel's second pair is undenominated
        recs, keep = tr.trace_tuple(tup, 8)
        z0, r8, l8, f8, two = evaluate(recs)
        assert int(recs[-1].K) == 3 and two and not np.any(z0) and np.array_equal(r8, rt) and np.array_equal(l8, lt) and f8 == f
        eps8 = _eps_of(recs[-1], rt.shape)
        assert all(np.array_equal(eps8[i], eps[i]) for i in range(6))
        a8, apc8 = _energies(t, tup, r8, r8, l8, f, eps8)
        b8, bpd8 = f * np.sum(l8 * r8), f * np.sum(l8 * (r8 + l8))
        got8 = np.array([a8, b8, apc8 - a8, bpd8 - b8])
        assert np.max(np.abs(got8 - sums_ref)) <= 1e-13 * max(1.0, np.max(np.abs(sums_ref)))
        n += 1
    tr.close()
    assert n >= 5


def test_trace_context_error_paths():
    """Status + message instead of a crash: intermediates from another tiling, CR-EOM without the ground-state
    intermediates, x amplitudes with a foreign block structure, a method the context was not set up for."""
    from oracle import cr_dense
    t = tl.make_tiling(OCC, VIRT, 2)
    t_other = tl.make_tiling([3, 1], [4, 2], 2)
    st = synth.physical(t)
    d = cr_dense.DenseEOM(t)
    cr, q = d.stores()
    cr_other = cr_dense.Dense(t_other).stores()
    tr = capi.Triples(trace=True)
    tr.set_state(st)
    tup = [int(x) for x in oracle_task(t)]
    with pytest.raises(RuntimeError):
        tr.trace_tuple(tup, 2)                       # CR before set_cr
    with pytest.raises(RuntimeError):
        tr.set_creom(q)                              # r0 != 0 needs set_cr
    with pytest.raises(RuntimeError):
        tr.set_cr(cr_other)                          # offset tables of another tiling
    tr.set_cr(cr)
    bad = dataclasses.replace(q, x2_hash=cr.n1_hash)
    with pytest.raises(RuntimeError):
        tr.set_creom(bad)
    tr.set_creom(q)
    recs, keep = tr.trace_tuple(tup, 8)
    assert recs[-1].kind == 9
    with pytest.raises(RuntimeError):
        tr.trace_tuple(tup, 99)
    with pytest.raises(RuntimeError):
        tr.run_cr()                                  # a trace context cannot compute
    tr.close()


def oracle_task(t):
    from oracle import oracle as ora
    return ora.task_list(t)[0][:6]
