"""ctypes binding of oracle/liboracle_triples.so -- TEST INFRASTRUCTURE ONLY (see triples_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_triples.so")
L = C.c_long
PD = C.POINTER(C.c_double)
PL = C.POINTER(C.c_long)


def build(quiet=True):
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


class Ctx(C.Structure):
    _fields_ = [("noab", L), ("nvab", L), ("restricted", L), ("irrep_t", L), ("irrep_v", L),
                ("spin", PL), ("sym", PL), ("range", PL), ("offset", PL), ("alpha", PL), ("evl_sorted", PD),
                ("t1_hash", PL), ("t1", PD), ("t2_hash", PL), ("t2", PD), ("v2_hash", PL), ("v2", PD),
                ("intorb", L), ("noa", L), ("nva", L), ("b2am", PL), ("spin_alpha", PL), ("sym_alpha", PL),
                ("range_alpha", PL), ("v2orb_hash", PL), ("v2orb", PD)]


class Counts(C.Structure):
    _fields_ = [("flops_s1", C.c_double), ("flops_d1", C.c_double), ("flops_d2", C.c_double),
                ("calls_s1", L), ("calls_d1", L), ("calls_d2", L)]

    @property
    def flops(self):
        return self.flops_s1 + self.flops_d1 + self.flops_d2


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.ora_ccsd_t.restype = L
        _lib.ora_ccsd_t_6tasks.restype = L
        _lib.ora_tce_hash.restype = L
        _lib.ora_tce_tile_group.restype = L
        _lib.ora_ccsd_t_factor.restype = C.c_double
    return _lib


def _pl(a):
    return a.ctypes.data_as(PL)


def _pd(a):
    return a.ctypes.data_as(PD)


def make_ctx(st):
    """st: nwchem_b200.synth.BlockStores.  Returns (Ctx, keepalive)."""
    t = st.t
    arrs = dict(spin=np.ascontiguousarray(t.spin, np.int64), sym=np.ascontiguousarray(t.sym, np.int64),
                range=np.ascontiguousarray(t.range, np.int64), offset=np.ascontiguousarray(t.offset, np.int64),
                alpha=np.ascontiguousarray(t.alpha, np.int64), evl=np.ascontiguousarray(t.evl_sorted, np.float64),
                t1h=np.ascontiguousarray(st.t1_hash, np.int64), t1=np.ascontiguousarray(st.t1, np.float64),
                t2h=np.ascontiguousarray(st.t2_hash, np.int64), t2=np.ascontiguousarray(st.t2, np.float64),
                v2h=np.ascontiguousarray(st.v2_hash, np.int64), v2=np.ascontiguousarray(st.v2, np.float64))
    c = Ctx(t.noab, t.nvab, int(t.restricted), 0, 0, _pl(arrs["spin"]), _pl(arrs["sym"]), _pl(arrs["range"]),
            _pl(arrs["offset"]), _pl(arrs["alpha"]), _pd(arrs["evl"]), _pl(arrs["t1h"]), _pd(arrs["t1"]),
            _pl(arrs["t2h"]), _pd(arrs["t2"]), _pl(arrs["v2h"]), _pd(arrs["v2"]))
    orb = getattr(st, "orb", None)
    if orb is not None:   # `2eorb` storage: V2 comes from the orbital-form store (synth.OrbitalV2)
        a = orb.a
        arrs.update(b2am=np.ascontiguousarray(a.b2am, np.int64), spa=np.ascontiguousarray(a.spin_alpha, np.int64),
                    sya=np.ascontiguousarray(a.sym_alpha, np.int64), rga=np.ascontiguousarray(a.range_alpha, np.int64),
                    voh=np.ascontiguousarray(orb.v2orb_hash, np.int64), vo=np.ascontiguousarray(orb.v2orb, np.float64))
        c.intorb = 1; c.noa = a.noa; c.nva = a.nva
        c.b2am = _pl(arrs["b2am"]); c.spin_alpha = _pl(arrs["spa"]); c.sym_alpha = _pl(arrs["sya"])
        c.range_alpha = _pl(arrs["rga"]); c.v2orb_hash = _pl(arrs["voh"]); c.v2orb = _pd(arrs["vo"])
    return c, arrs


def task_list(t):
    l = lib()
    spin = np.ascontiguousarray(t.spin, np.int64); sym = np.ascontiguousarray(t.sym, np.int64)
    rng = np.ascontiguousarray(t.range, np.int64)
    n = l.ora_ccsd_t_6tasks(L(int(t.restricted)), L(t.noab), L(t.nvab), _pl(spin), _pl(sym))
    kl = np.zeros((max(n, 1), 7), np.int64)
    l.ora_ccsd_t_neword(L(n), L(int(t.restricted)), L(t.noab), L(t.nvab), _pl(spin), _pl(sym), _pl(rng), _pl(kl))
    return kl[:n]


def ccsd_t(st, per_task=False, count=False):
    """Whole (T) on the CPU: returns dict(e1, e2, tasks[, per_task, counts])."""
    l = lib()
    c, keep = make_ctx(st)
    n = len(task_list(st.t))
    e = np.zeros(2)
    kl = np.zeros((max(n, 1), 7), np.int64)
    pt = np.zeros((max(n, 1), 2))
    cnt = Counts()
    l.ora_ccsd_t(C.byref(c), _pd(e), _pl(kl), _pd(pt), C.byref(cnt))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return dict(e1=float(e[0]), e2=float(e[1]), tasks=kl[:n], per_task=pt[:n], counts=cnt)


def ccsd_t_restart(st, begin=1, table=None, max_outer=0):
    """Restartable (T), ccsd_t_restart.F: returns (new begin, table[nvab], t_energy, outer tiles done)."""
    l = lib()
    c, keep = make_ctx(st)
    tab = np.zeros(st.t.nvab) if table is None else np.ascontiguousarray(table, np.float64).copy()
    b = L(begin)
    te = C.c_double(0.0)
    l.ora_ccsd_t_restart.restype = L
    done = l.ora_ccsd_t_restart(C.byref(c), C.byref(b), _pd(tab), L(max_outer), C.byref(te))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return int(b.value), tab, float(te.value), int(done)


def hash_v2(orb, key, irrep_v=0):
    """tce_hash_v2 on an orbital-form store (synth.OrbitalV2): offset of block `key` or -1."""
    l = lib()
    a = orb.a
    l.ora_tce_hash_v2.restype = L
    h = np.ascontiguousarray(orb.v2orb_hash, np.int64)
    sp = np.ascontiguousarray(a.spin_alpha, np.int64); sy = np.ascontiguousarray(a.sym_alpha, np.int64)
    rg = np.ascontiguousarray(a.range_alpha, np.int64)
    return int(l.ora_tce_hash_v2(_pl(h), L(key), L(a.noa), L(a.nva), _pl(sp), _pl(sy), _pl(rg), L(irrep_v)))


def v2_block_intorb(st, g3b, g4b, g1b, g2b):
    """get_block_ind_i: the antisymmetrised spin-orbital block <g3b g4b||g1b g2b> built from the orbital store."""
    l = lib()
    c, keep = make_ctx(st)
    n = st.t.r(g3b) * st.t.r(g4b) * st.t.r(g1b) * st.t.r(g2b)
    out = np.zeros(n)
    l.ora_get_block_ind_i(C.byref(c), _pd(out), L(n), L(g2b), L(g1b), L(g4b), L(g3b))
    if l.ora_error():
        raise RuntimeError("oracle: orbital block not found")
    return out


def singles_tce(st, tup):
    """Singles tile of one tuple through the original TCE formulation (ccsd_t_singles.F: sort, outer product,
    nine TCE_SORTACC_6), indexed [p4,p5,p6,h1,h2,h3]."""
    l = lib()
    c, keep = make_ctx(st)
    dims = [st.t.r(b) for b in tup]
    s = np.zeros(int(np.prod(dims)))
    p4, p5, p6, h1, h2, h3 = [L(int(x)) for x in tup]
    l.ora_ccsd_t_singles_tce(C.byref(c), _pd(s), h1, h2, h3, p4, p5, p6)
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return s.reshape(dims)


def doubles_tce(st, tup):
    """Doubles tile of one tuple through the original TCE formulation (ccsd_t_doubles.F: sorts, DGEMM, eighteen
    TCE_SORTACC_6), indexed [p4,p5,p6,h1,h2,h3]."""
    l = lib()
    c, keep = make_ctx(st)
    dims = [st.t.r(b) for b in tup]
    d = np.zeros(int(np.prod(dims)))
    p4, p5, p6, h1, h2, h3 = [L(int(x)) for x in tup]
    l.ora_ccsd_t_doubles_tce(C.byref(c), _pd(d), h1, h2, h3, p4, p5, p6)
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return d.reshape(dims)


def tuple_tiles(st, tup):
    """One tuple (p4b,p5b,p6b,h1b,h2b,h3b): returns (singles, doubles, e1, e2) with the t3 tiles as
    arrays indexed [p4,p5,p6,h1,h2,h3] (C order == Fortran T3(h3,h2,h1,p6,p5,p4))."""
    l = lib()
    c, keep = make_ctx(st)
    t = st.t
    dims = [t.r(b) for b in tup]
    size = int(np.prod(dims))
    s = np.zeros(size); d = np.zeros(size); e = np.zeros(2)
    tt = np.array(tup, np.int64)
    cnt = Counts()
    l.ora_ccsd_t_loop(C.byref(c), _pl(tt), _pd(s), _pd(d), _pd(e), C.byref(cnt))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return s.reshape(dims), d.reshape(dims), float(e[0]), float(e[1]), cnt


def tuple_slab(st, tup, lo, hi, tiles=False):
    """One tuple restricted to the p4 slab [lo,hi) of its t3 tile (the slicing of ccsd_t_6dts.F restated on the 27
    kernels): returns (e1, e2) of the slab [, singles, doubles indexed [p4-lo,p5,p6,h1,h2,h3]].  Slabs add up to the tuple."""
    l = lib()
    c, keep = make_ctx(st)
    t = st.t
    hi = min(hi, t.r(int(tup[0])))
    dims = [hi - lo] + [t.r(int(b)) for b in tup[1:6]]
    size = int(np.prod(dims))
    s = np.zeros(size); d = np.zeros(size); e = np.zeros(2)
    tt = np.array(tup[:6], np.int64)
    l.ora_ccsd_t_loop_slab(C.byref(c), _pl(tt), L(lo), L(hi), _pd(s), _pd(d), _pd(e))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    if tiles:
        return float(e[0]), float(e[1]), s.reshape(dims), d.reshape(dims)
    return float(e[0]), float(e[1])


def tuple_sliced(st, tup, width=4):
    """Whole tuple, p4-sliced `width` at a time (never holds more than width/range(p4) of the tile)."""
    e1 = e2 = 0.0
    for lo in range(0, st.t.r(int(tup[0])), width):
        a, b = tuple_slab(st, tup, lo, lo + width)
        e1 += a; e2 += b
    return e1, e2


class Lambda(C.Structure):   # ora_lambda
    _fields_ = [("y1_hash", PL), ("y1", PD), ("y2_hash", PL), ("y2", PD), ("f1_hash", PL), ("f1", PD),
                ("irrep_y", L), ("irrep_f", L)]


def make_lambda(lam):
    k = dict(y1h=np.ascontiguousarray(lam.y1_hash, np.int64), y1=np.ascontiguousarray(lam.y1, np.float64),
             y2h=np.ascontiguousarray(lam.y2_hash, np.int64), y2=np.ascontiguousarray(lam.y2, np.float64),
             f1h=np.ascontiguousarray(lam.f1_hash, np.int64), f1=np.ascontiguousarray(lam.f1, np.float64))
    return Lambda(_pl(k["y1h"]), _pd(k["y1"]), _pl(k["y2h"]), _pd(k["y2"]), _pl(k["f1h"]), _pd(k["f1"]), 0, 0), k


def lambda_ccsd_t(st, lam, sorted=True):
    """Lambda-CCSD(T) on the CPU (lambda_ccsd_t.F + lambda_ccsd_t_left.F restated).  sorted=False reproduces the file
    literally (L3-ordered left tiles multiplied element-wise with the T3-ordered right tile); sorted=True applies the
    sort its declarations announce.  Returns dict(e1, e2, per_task) in the loop order of lambda_ccsd_t.F:59-64."""
    l = lib()
    c, keep = make_ctx(st)
    y, keep2 = make_lambda(lam)
    n = len(task_list(st.t))
    e = np.zeros(2); pt = np.zeros((max(n, 1), 2))
    l.ora_lambda_ccsd_t.restype = L
    cnt = l.ora_lambda_ccsd_t(C.byref(c), C.byref(y), int(bool(sorted)), _pd(e), _pd(pt))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return dict(e1=float(e[0]), e2=float(e[1]), per_task=pt[:cnt])


def lambda_tuple(st, lam, tup, sorted=True):
    """One tuple (p4b..h3b): (e1, e2, tdoubles [p4,p5,p6,h1,h2,h3], ysingles, ydoubles [h1,h2,h3,p4,p5,p6])."""
    l = lib()
    c, keep = make_ctx(st)
    y, keep2 = make_lambda(lam)
    t = st.t
    dims = [t.r(int(b)) for b in tup[:6]]
    n = int(np.prod(dims))
    td = np.zeros(n); ys = np.zeros(n); yd = np.zeros(n); e = np.zeros(2)
    tt = np.array(tup[:6], np.int64)
    l.ora_lambda_ccsd_t_tuple(C.byref(c), C.byref(y), _pl(tt), int(bool(sorted)), _pd(e), _pd(td), _pd(ys), _pd(yd))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    ld = dims[3:] + dims[:3]
    return float(e[0]), float(e[1]), td.reshape(dims), ys.reshape(ld), yd.reshape(ld)


def count_tuple(st_or_ctx, tup, keep=None):
    l = lib()
    c, keep = make_ctx(st_or_ctx) if keep is None else (st_or_ctx, keep)
    cnt = Counts()
    tt = np.array(tup, np.int64)
    l.ora_ccsd_t_count(C.byref(c), _pl(tt), C.byref(cnt))
    return cnt


def kernel(family, k, dims_perm, kd, triplesx, tsub, v2sub):
    """Call one of the 27 CPU kernels. dims_perm = (h3d,h2d,h1d,p6d,p5d,p4d) of the PERMUTED tuple."""
    l = lib()
    h3d, h2d, h1d, p6d, p5d, p4d = [int(x) for x in dims_perm]
    l.ora_sd_t_kernel(L(family), L(k), L(h3d), L(h2d), L(h1d), L(p6d), L(p5d), L(p4d), L(int(kd)),
                      _pd(triplesx), _pd(tsub), _pd(v2sub))


def kernel_slab(family, k, dims_perm, kd, lo, hi, triplesx, tsub, v2sub):
    """One of the 27 CPU kernels restricted to the p4 slab [lo,hi) of the TASK tuple's tile; triplesx holds the slab."""
    l = lib()
    h3d, h2d, h1d, p6d, p5d, p4d = [int(x) for x in dims_perm]
    l.ora_sd_t_kernel_slab(L(family), L(k), L(h3d), L(h2d), L(h1d), L(p6d), L(p5d), L(p4d), L(int(kd)), L(lo), L(hi),
                           _pd(triplesx), _pd(tsub), _pd(v2sub))


def set_num_threads(n):
    lib().ora_set_num_threads(int(n))


def tile_group(n, isize):
    l = lib()
    out = np.zeros(max(n, 1), np.int64)
    k = l.ora_tce_tile_group(L(n), L(isize), _pl(out))
    return [int(x) for x in out[:k]]


def num_threads():
    return int(lib().ora_num_threads())


class CR(C.Structure):   # ora_cr (cr_oracle.h)
    _fields_ = [("n1_hash", PL), ("n1", PD), ("n2_hash", PL), ("n2", PD), ("e2_hash", PL), ("e2", PD)]


def make_cr(cr):
    k = dict(n1h=np.ascontiguousarray(cr.n1_hash, np.int64), n1=np.ascontiguousarray(cr.n1, np.float64),
             n2h=np.ascontiguousarray(cr.n2_hash, np.int64), n2=np.ascontiguousarray(cr.n2, np.float64),
             e2h=np.ascontiguousarray(cr.e2_hash, np.int64), e2=np.ascontiguousarray(cr.e2, np.float64))
    return CR(_pl(k["n1h"]), _pd(k["n1"]), _pl(k["n2h"]), _pd(k["n2"]), _pl(k["e2h"]), _pd(k["e2"])), k


def cr_ccsd_t(st, cr):
    """CR-CCSD(T) tuple loop on the CPU (cr_ccsd_t.F:93-263 restated, cr_oracle.h).  `cr` = cr_dense.CRStores (the three
    intermediates + den0).  Returns dict(sums = (num1,num2,den1,den2) without den0, per_task[n,4] in the loop order of
    cr_ccsd_t.F:93-98, tasks[n,6], e1, e2 = the CR-CCSD[T] / CR-CCSD(T) corrections :262-263)."""
    l = lib()
    c, keep = make_ctx(st)
    y, keep2 = make_cr(cr)
    n = len(task_list(st.t))
    s = np.zeros(4); pt = np.zeros((max(n, 1), 4))
    l.ora_cr_ccsd_t.restype = L
    cnt = l.ora_cr_ccsd_t(C.byref(c), C.byref(y), _pd(s), _pd(pt))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    den0 = float(getattr(cr, "den0", 0.0))
    return dict(sums=s.copy(), per_task=pt[:cnt], e1=float(s[0] / (1.0 + s[2] + den0)), e2=float(s[1] / (1.0 + s[3] + den0)))


def lr_ccsd_t(st, cr):
    """LR-CCSD(T) tuple loop on the CPU (lr_ccsd_t.F restated, cr_oracle.h): the six corrections (IA, IB, IIA, IIB, IIIA,
    IIIB) of lr_ccsd_t.F:408-413.  Test infrastructure for the QA golden numbers of tce_lr_ccsd_t only."""
    l = lib()
    c, keep = make_ctx(st)
    y, keep2 = make_cr(cr)
    s = np.zeros(6)
    l.ora_lr_ccsd_t.restype = L
    l.ora_lr_ccsd_t(C.byref(c), C.byref(y), _pd(s))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return dict(zip(("IA", "IB", "IIA", "IIB", "IIIA", "IIIB"), (float(x) for x in s)))


def cr_tuple(st, cr, tup):
    """One tuple: (sums[4], moment tile, denominator tile), tiles indexed [p4,p5,p6,h1,h2,h3]."""
    l = lib()
    c, keep = make_ctx(st)
    y, keep2 = make_cr(cr)
    dims = [st.t.r(int(b)) for b in tup[:6]]
    n = int(np.prod(dims))
    s = np.zeros(4); m = np.zeros(n); e = np.zeros(n)
    tt = np.array(tup[:6], np.int64)
    l.ora_cr_ccsd_t_tuple(C.byref(c), C.byref(y), _pl(tt), _pd(s), _pd(m), _pd(e))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return s, m.reshape(dims), e.reshape(dims)


class CREOM(C.Structure):   # ora_creom (cr_oracle.h)
    _fields_ = [("x1_hash", PL), ("x1", PD), ("x2_hash", PL), ("x2", PD), ("m1_hash", PL), ("m1", PD), ("m2_hash", PL), ("m2", PD),
                ("m3_hash", PL), ("m3", PD), ("m4_hash", PL), ("m4", PD), ("q2_hash", PL), ("q2", PD),
                ("r0", C.c_double), ("excit", C.c_double), ("lr0", L)]


def make_creom(q):
    names = ("x1", "x2", "m1", "m2", "m3", "m4", "q2")
    k = {}
    args = []
    for n in names:
        k[n + "h"] = np.ascontiguousarray(getattr(q, n + "_hash"), np.int64)
        k[n] = np.ascontiguousarray(getattr(q, n), np.float64)
        args += [_pl(k[n + "h"]), _pd(k[n])]
    return CREOM(*args, float(q.r0), float(q.excit), int(abs(q.r0) >= 1e-7)), k


def cr_eomccsd_t(st, cr, q):
    """CR-EOMCCSD(T) tuple loop on the CPU (cr_eomccsd_t.F:325-493 restated, cr_oracle.h).  cr = cr_dense.CRStores (the
    ground-state intermediates, read when r0 != 0), q = cr_dense.CREOMStores.  Returns dict(sums = (sum f R R/denex,
    sum f L R, sum f L R/denex, sum f L L), per_task[n,4], num1, den1)."""
    l = lib()
    c, keep = make_ctx(st)
    y, keep2 = make_cr(cr)
    z, keep3 = make_creom(q)
    n = len(task_list(st.t))
    s = np.zeros(4); pt = np.zeros((max(n, 1), 4))
    l.ora_cr_eomccsd_t.restype = L
    cnt = l.ora_cr_eomccsd_t(C.byref(c), C.byref(y), C.byref(z), _pd(s), _pd(pt))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return dict(sums=s.copy(), per_task=pt[:cnt], num1=float(s[0] + s[1]), den1=float(s[2] + s[3]))


def cr_eom_tuple(st, cr, q, tup):
    """One tuple: (sums[4], right tile, left tile), tiles indexed [p4,p5,p6,h1,h2,h3]."""
    l = lib()
    c, keep = make_ctx(st)
    y, keep2 = make_cr(cr)
    z, keep3 = make_creom(q)
    dims = [st.t.r(int(b)) for b in tup[:6]]
    n = int(np.prod(dims))
    s = np.zeros(4); r = np.zeros(n); le = np.zeros(n)
    tt = np.array(tup[:6], np.int64)
    l.ora_cr_eomccsd_t_tuple(C.byref(c), C.byref(y), C.byref(z), _pl(tt), _pd(s), _pd(r), _pd(le))
    if l.ora_error():
        raise RuntimeError("oracle: block key not found")
    return s, r.reshape(dims), le.reshape(dims)


def cre_t_vs_d2cp(k0, dims_perm, kd, t2sub, v2sub):
    """(cre_t_K with the factor of creomsd_t_n2_mem_2, with the factor of _4, sd_t_d2cp_K) on the same operands."""
    l = lib()
    h3d, h2d, h1d, p6d, p5d, p4d = [int(x) for x in dims_perm]
    n = h3d * h2d * h1d * p6d * p5d * p4d
    a = np.zeros(n); b = np.zeros(n); d = np.zeros(n)
    l.ora_cre_t_vs_d2cp(L(k0), L(h3d), L(h2d), L(h1d), L(p6d), L(p5d), L(p4d), L(int(kd)), _pd(t2sub), _pd(v2sub), _pd(a), _pd(b), _pd(d))
    return a, b, d
