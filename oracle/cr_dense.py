"""CR-CCSD(T) intermediates and an untiled reference of the whole correction -- TEST INFRASTRUCTURE ONLY.

The per-tuple half of CR-CCSD(T) (cr_ccsd_t.F:93-233: cr_ccsd_t_N_1 / _N_2 / _E_1 / _E_2 and the four energy sums) is
restated line by line in triples_oracle.c.  Its inputs are three intermediate block stores the reference builds once,
before the tuple loop, with ~25 TCE-generated block contraction routines (cr_ccsd_t_N.F:665-6200 with toggle 1,
cr_ccsd_t_E.F:743-905) -- or loads from files (read_in3: gr1_1, gr1_2, ei1_2; cr_ccsd_t_N.F:98-104).  Those routines are
upstream of the hot path (SURVEY 8 f4 territory); here the tensors they produce are evaluated DENSELY in the spin-orbital
basis straight from the tensor-contraction expressions the TCE printed at the top of each file (the specification the
generated Fortran implements; cited per term below), and packed into the reference's block layouts
(nwchem_b200.tiling.cr_n1_offset / cr_n2_offset / cr_e2_offset).  Only for small orbital counts.

`dense_reference` evaluates the four sums and den0 of cr_ccsd_t.F without any tiling at all (full antisymmetric
tensors, unrestricted sums divided by 36): the check that the tiled restatement -- permutation tables, dispatch tests,
restricted mapping, kernels, factors -- and the packing agree with the algebra."""
from __future__ import annotations
import dataclasses
import numpy as np
from nwchem_b200 import synth, tiling as tl


@dataclasses.dataclass
class CRStores:
    """The three intermediates cr_ccsd_t.F holds during the tuple loop + the scalar of cr_ccsd_t_D."""
    n1_hash: np.ndarray; n1: np.ndarray     # d_i1_1: i1(h11 p4 h1 h2), cr_ccsd_t_N.F:9-25
    n2_hash: np.ndarray; n2: np.ndarray     # d_i1_2: i1(p4 p5 h1 p12), cr_ccsd_t_N.F:27-39
    e2_hash: np.ndarray; e2: np.ndarray     # d_i1_3: i1(p4 p5 h1 h2)_tt, cr_ccsd_t_E.F:9
    den0: float                             # cr_ccsd_t_D.F:6-9


class Dense:
    """Dense spin-orbital tensors of the closed-shell synthetic problem of synth.physical.  Spin-orbital g of spatial
    orbital x and spin s (0 alpha, 1 beta) is x + n*s; holes are the x < no, particles the x >= no."""

    def __init__(self, t: tl.Tiling, seed: int = 20240229, fock_seed: int = 4242, t_scale: float = 1.0, dense=None,
                 fock_hp=None):
        """dense = (no, nv, t1s, t2s, eri) replaces the synthetic tensors (real amplitudes: oracle/h2o_ccsd.py), in the
        spatial-orbital order of the tiling; fock_hp[i,a] then gives the (hole, particle) Fock block (zero for canonical
        Hartree-Fock orbitals)."""
        no, nv, t1s, t2s, eri = synth.physical_dense(t, seed) if dense is None else dense
        t1s = t1s * t_scale; t2s = t2s * t_scale
        n = no + nv
        self.t, self.no, self.nv, self.n = t, no, nv, n
        self.H = np.array([i + n * s for s in (0, 1) for i in range(no)])
        self.P = np.array([no + a + n * s for s in (0, 1) for a in range(nv)])
        spin = np.repeat([0, 1], n)
        spat = np.tile(np.arange(n), 2)
        same = (spin[:, None] == spin[None, :]).astype(float)
        # <pq||rs> = (pr|qs) d(sp,sr) d(sq,ss) - (ps|qr) d(sp,ss) d(sq,sr)
        e = eri[np.ix_(spat, spat, spat, spat)]
        self.v = (np.einsum("prqs,pr,qs->pqrs", e, same, same) - np.einsum("psqr,ps,qr->pqrs", e, same, same))
        sp_p, sp_h = spin[self.P], spin[self.H]
        xa, xi = spat[self.P] - no, spat[self.H]
        d_ph = (sp_p[:, None] == sp_h[None, :]).astype(float)
        self.t1 = t1s[np.ix_(xa, xi)] * d_ph                                  # t(p,h)
        tt = t2s[np.ix_(xa, xa, xi, xi)]
        self.t2 = (tt * d_ph[:, None, :, None] * d_ph[None, :, None, :]
                   - tt.transpose(0, 1, 3, 2) * d_ph[:, None, None, :] * d_ph[None, :, :, None])   # t(p,p,h,h)
        # Fock (hole, particle) block: nonzero for a non-canonical reference; same irrep mask as t1
        rng = np.random.default_rng(fock_seed)
        fs = rng.uniform(-1, 1, (no, nv)) * 0.01
        irr = np.zeros(n, dtype=np.int64)
        for b in range(t.noab + t.nvab):
            if t.spin[b] == 1:
                irr[t.members[b]] = t.sym[b]
        fs[(irr[:no, None] ^ irr[None, no:]) != 0] = 0.0
        if dense is not None:
            fs = np.zeros((no, nv)) if fock_hp is None else np.asarray(fock_hp, dtype=np.float64)
        self.fs = fs
        self.f_hp = fs[np.ix_(xi, xa)] * d_ph.T                                # f(h,p)
        eps = np.zeros(2 * n)
        for b in range(t.noab + t.nvab):
            s = int(t.spin[b]) - 1
            eps[t.members[b] + n * s] = t.evl_sorted[t.offset[b]:t.offset[b] + t.range[b]]
        self.eps = eps
        H, P = self.H, self.P
        v = self.v
        self.v_hhhh = v[np.ix_(H, H, H, H)]; self.v_hhhp = v[np.ix_(H, H, H, P)]; self.v_hhpp = v[np.ix_(H, H, P, P)]
        self.v_hphh = v[np.ix_(H, P, H, H)]; self.v_hphp = v[np.ix_(H, P, H, P)]; self.v_hppp = v[np.ix_(H, P, P, P)]
        self.v_pphp = v[np.ix_(P, P, H, P)]; self.v_pppp = v[np.ix_(P, P, P, P)]; self.v_pphh = v[np.ix_(P, P, H, H)]

    # ---- cr_ccsd_t_N.F:9-25: i1(h11 p4 h1 h2), indexed [h11,p4,h1,h2] ----
    def n1(self):
        t1, t2 = self.t1, self.t2
        E = np.einsum
        # :13-14  i3(h7 h11 h1 p8) = v(h7 h11 h1 p8) - 1/2 Sum(p9) t(p9 h1) v(h7 h11 p8 p9)
        i3 = self.v_hhhp - 0.5 * E("ei,abce->abic", t1, self.v_hhpp)
        # :11-15  i2(h7 h11 h1 h2) = v - P(2) Sum(p8) t(p8 h1) i3(h7 h11 h2 p8) + 1/2 Sum(p8 p9) t(p8 p9 h1 h2) v(h7 h11 p8 p9)
        x = E("ei,abje->abij", t1, i3)
        i2a = self.v_hhhh - (x - x.transpose(0, 1, 3, 2)) + 0.5 * E("efij,abef->abij", t2, self.v_hhpp)
        # :17-18  i2(h11 p4 h1 p9) = v(h11 p4 h1 p9) + 1/2 Sum(p8) t(p8 h1) v(h11 p4 p8 p9)
        i2b = self.v_hphp + 0.5 * E("ei,bcef->bcif", t1, self.v_hppp)
        # :20-21  i2(h11 p12) = f(h11 p12) + Sum(h10 p9) t(p9 h10) v(h10 h11 p9 p12)
        i2c = self.f_hp + E("em,mbef->bf", t1, self.v_hhpp)
        # :23-24  i2(h9 h11 h1 p8) = v(h9 h11 h1 p8) - Sum(p10) t(p10 h1) v(h9 h11 p8 p10)
        i2d = self.v_hhhp - E("fi,abef->abie", t1, self.v_hhpp)
        out = self.v_hphh.copy()                                              # :10
        out += E("cm,mbij->bcij", t1, i2a)                                    # :11  + Sum(h7) t(p4 h7) i2(h7 h11 h1 h2)
        x = E("ei,bcje->bcij", t1, i2b)                                       # :16  - P(2) Sum(p9) t(p9 h1) i2(h11 p4 h2 p9)
        out -= x - x.transpose(0, 1, 3, 2)
        out -= E("ceij,be->bcij", t2, i2c)                                    # :19  - Sum(p12) t(p4 p12 h1 h2) i2(h11 p12)
        x = E("ceim,mbje->bcij", t2, i2d)                                     # :22  + P(2) Sum(h9 p8) t(p4 p8 h1 h9) i2(h9 h11 h2 p8)
        out += x - x.transpose(0, 1, 3, 2)
        out += 0.5 * E("efij,bcef->bcij", t2, self.v_hppp)                    # :25  + 1/2 Sum(p8 p9) t(p8 p9 h1 h2) v(h11 p4 p8 p9)
        return out

    # ---- cr_ccsd_t_N.F:27-39: i1(p4 p5 h1 p12), indexed [p4,p5,h1,p12] ----
    def n2(self):
        t1, t2 = self.t1, self.t2
        E = np.einsum
        # :31-32  i3(h8 h11 h1 p12) = v(h8 h11 h1 p12) + Sum(p9) t(p9 h1) v(h8 h11 p9 p12)
        j3 = self.v_hhhp + E("ei,abef->abif", t1, self.v_hhpp)
        # :29-34  i2(h11 p4 h1 p12) = v + 1/2 Sum(h8) t(p4 h8) i3(h8 h11 h1 p12) + Sum(p8) t(p8 h1) v(h11 p4 p8 p12)
        #                             - Sum(h9 p8) t(p4 p8 h1 h9) v(h9 h11 p8 p12)
        j2a = (self.v_hphp + 0.5 * E("cm,mbif->bcif", t1, j3) + E("ei,bcef->bcif", t1, self.v_hppp)
               - E("ceim,mbef->bcif", t2, self.v_hhpp))
        # :37-38  i2(h8 h9 h1 p12) = v(h8 h9 h1 p12) + Sum(p10) t(p10 h1) v(h8 h9 p10 p12)
        j2b = self.v_hhhp + E("ei,abef->abif", t1, self.v_hhpp)
        out = self.v_pphp.copy()                                              # :28
        x = E("cm,mdif->cdif", t1, j2a)                                       # :29  - P(2) Sum(h11) t(p4 h11) i2(h11 p5 h1 p12)
        out -= x - x.transpose(1, 0, 2, 3)
        out += E("ei,cdef->cdif", t1, self.v_pppp)                            # :35  + Sum(p8) t(p8 h1) v(p4 p5 p8 p12)
        out += 0.5 * E("cdmn,mnif->cdif", t2, j2b)                            # :36  + 1/2 Sum(h8 h9) t(p4 p5 h8 h9) i2(h8 h9 h1 p12)
        x = E("ceim,mdef->cdif", t2, self.v_hppp)                             # :39  + P(2) Sum(h9 p8) t(p4 p8 h1 h9) v(h9 p5 p8 p12)
        out += x - x.transpose(1, 0, 2, 3)
        return out

    # ---- cr_ccsd_t_E.F:9: i1(p4 p5 h1 h2) = -1/4 P(4) t(p4 h1) t(p5 h2), indexed [p4,p5,h1,h2] ----
    def e2(self):
        x = np.einsum("ai,bj->abij", self.t1, self.t1)
        return -0.25 * (x - x.transpose(1, 0, 2, 3) - x.transpose(0, 1, 3, 2) + x.transpose(1, 0, 3, 2))

    # ---- cr_ccsd_t_D.F:6-9 with c = t (cr_ccsd_t.F:66-67) ----
    def den0(self):
        t1, t2 = self.t1, self.t2
        i1 = t1.T + 0.5 * np.einsum("cami,cm->ia", t2, t1)       # i1(h6 p5) = c+(h6 p5) + 1/2 Sum(h4 p3) c+(h4 h6 p3 p5) t(p3 h4)
        return float(np.einsum("ai,ia->", t1, i1) + 0.25 * np.einsum("abij,abij->", t2, t2))

    # ---- the four six-index tensors of cr_ccsd_t.F:139-152, indexed [p4,p5,p6,h1,h2,h3], fully antisymmetric ----
    @staticmethod
    def _p9(x, lone_p, lone_h):
        """Antisymmetriser P(9) of a term whose particle `lone_p` (0,1,2) and hole `lone_h` sit on the other factor."""
        def sw(a, i, j, off):
            ax = list(range(6))
            ax[off + i], ax[off + j] = ax[off + j], ax[off + i]
            return a.transpose(ax)
        others_p = [i for i in range(3) if i != lone_p]
        others_h = [i for i in range(3) if i != lone_h]
        y = x - sw(x, lone_p, others_p[0], 0) - sw(x, lone_p, others_p[1], 0)
        return y - sw(y, lone_h, others_h[0], 3) - sw(y, lone_h, others_h[1], 3)

    def six_index(self, n1=None, n2=None):
        t1, t2 = self.t1, self.t2
        E = np.einsum
        n1 = self.n1() if n1 is None else n1
        n2 = self.n2() if n2 is None else n2
        # ccsd_t_singles.F:6  P(9) t(p4 h1) v(p5 p6 h2 h3)
        S = self._p9(E("ai,bcjk->abcijk", t1, self.v_pphh), 0, 0)
        # ccsd_t_doubles.F:37 / :288  -P(9) Sum(h7) t(p4 p5 h1 h7) v(h7 p6 h2 h3) - P(9) Sum(p7) t(p4 p7 h1 h2) v(p5 p6 h3 p7)
        D = -self._p9(E("abim,mcjk->abcijk", t2, self.v_hphh), 2, 0) - self._p9(E("aeij,bcke->abcijk", t2, self.v_pphp), 0, 2)
        # cr_ccsd_t_N.F:8 / :26  same with the dressed intermediates
        M = -self._p9(E("abim,mcjk->abcijk", t2, n1), 2, 0) - self._p9(E("aeij,bcke->abcijk", t2, n2), 0, 2)
        # cr_ccsd_t_E.F:7-8  P(9) t(p4 p5 h1 h2) t(p6 h3) - 2/3 P(9) t(p4 h1) i1(p5 p6 h2 h3)
        Et = self._p9(E("abij,ck->abcijk", t2, t1), 2, 2) - (2.0 / 3.0) * self._p9(E("ai,bcjk->abcijk", t1, self.e2()), 0, 0)
        return S, D, M, Et

    def dense_reference(self):
        """(num1, num2, den1, den2, den0) of cr_ccsd_t.F:176-205, :257-258 from the untiled tensors."""
        S, D, M, Et = self.six_index()
        ep, eh = self.eps[self.P], self.eps[self.H]
        delta = (-ep[:, None, None, None, None, None] - ep[None, :, None, None, None, None] - ep[None, None, :, None, None, None]
                 + eh[None, None, None, :, None, None] + eh[None, None, None, None, :, None] + eh[None, None, None, None, None, :])
        w = 1.0 / (36.0 * delta)
        return (float(np.sum(M * D * w)), float(np.sum(M * (S + D) * w)), float(np.sum(Et * D * w)),
                float(np.sum(Et * (S + D) * w)), self.den0())

    # ---- packing into the reference's block stores ----
    def _so(self, b):
        t = self.t
        return t.members[b - 1] + self.n * (int(t.spin[b - 1]) - 1)

    def _hidx(self, b):   # positions of tile b's orbitals inside self.H
        pos = {int(g): k for k, g in enumerate(self.H)}
        return np.array([pos[int(g)] for g in self._so(b)])

    def _pidx(self, b):
        pos = {int(g): k for k, g in enumerate(self.P)}
        return np.array([pos[int(g)] for g in self._so(b)])

    def stores(self) -> CRStores:
        t = self.t
        n1d, n2d, e2d = self.n1(), self.n2(), self.e2()
        h1, s1 = tl.cr_n1_offset(t); h2, s2 = tl.cr_n2_offset(t); h3, s3 = tl.cr_e2_offset(t)
        a1 = np.zeros(s1); a2 = np.zeros(s2); a3 = np.zeros(s3)
        for key, off in synth._iter_hash(h1):
            p4b, h11b, h1b, h2b = tl.decode_cr_n1_key(t, key)
            # stored (p4, h11, h1, h2), h2 fastest (the two TCE_SORT_4 of cr_ccsd_t_N_1_1, cr_ccsd_t_N.F:740-750)
            blk = n1d[np.ix_(self._hidx(h11b), self._pidx(p4b), self._hidx(h1b), self._hidx(h2b))].transpose(1, 0, 2, 3)
            a1[off:off + blk.size] = blk.ravel()
        for key, off in synth._iter_hash(h2):
            p4b, p5b, h1b, p12b = tl.decode_cr_n2_key(t, key)
            blk = n2d[np.ix_(self._pidx(p4b), self._pidx(p5b), self._hidx(h1b), self._pidx(p12b))]
            a2[off:off + blk.size] = blk.ravel()
        for key, off in synth._iter_hash(h3):
            p4b, p5b, h1b, h2b = tl.decode_t2_key(t, key)
            blk = e2d[np.ix_(self._pidx(p4b), self._pidx(p5b), self._hidx(h1b), self._hidx(h2b))]
            a3[off:off + blk.size] = blk.ravel()
        return CRStores(h1, a1, h2, a2, h3, a3, self.den0())

    def fock_store(self):
        """(f1_hash, f1): the (hole, particle) Fock blocks in the layout of tiling.f1_hp_offset (for callers that build
        the intermediates themselves; the per-tuple path does not read f)."""
        t = self.t
        fh, nf = tl.f1_hp_offset(t)
        f1 = np.zeros(nf)
        N = t.noab + t.nvab
        for key, off in synth._iter_hash(fh):
            h6b, p3b = key // N + 1, key % N + 1
            blk = self.f_hp[np.ix_(self._hidx(h6b), self._pidx(p3b))]
            f1[off:off + blk.size] = blk.ravel()
        return fh, f1


@dataclasses.dataclass
class CREOMStores:
    """Inputs of the CR-EOMCCSD(T) tuple loop beyond the ground-state ones (cr_eomccsd_t.F:262-288 builds them with toggle 1):
    the right-hand amplitudes x1 / x2, the four intermediates of creomsd_t_n2_mem, the one of q3rexpt2, r0 and the
    excitation energy."""
    x1_hash: np.ndarray; x1: np.ndarray
    x2_hash: np.ndarray; x2: np.ndarray
    m1_hash: np.ndarray; m1: np.ndarray     # d_i2_1  i1(h12 p4 h1 h2)
    m2_hash: np.ndarray; m2: np.ndarray     # d_i2_2  i1(p4 p5 h1 p12)
    m3_hash: np.ndarray; m3: np.ndarray     # d_i2_3  i1(h12 p4 h1 h2)
    m4_hash: np.ndarray; m4: np.ndarray     # d_i2_4  i1(p4 p5 h1 p10)
    q2_hash: np.ndarray; q2: np.ndarray     # d_i3_1  i1(p4 p5 h1 h2)_xt
    r0: float
    excit: float


class DenseEOM:
    """Synthetic inputs of the CR-EOMCCSD(T) tuple loop with the right index structure, and the untiled evaluation of the
    loop's four sums.  The EOM intermediates (creomccsd_t_n2_mem.F:9-90, ~80 TCE equations in f, v, t, x) are upstream of
    the hot path and are NOT restated: for the loop they are inputs like the amplitudes themselves, so any tensors with
    their symmetry do -- x1, x2 and the q3rexpt2 intermediate are the t1, t2, i1_tt of a second synthetic problem, the
    four moment intermediates the dressed hphh / pphp tensors of a second and third one."""

    def __init__(self, t: tl.Tiling, r0: float = 0.37, excit: float = 0.21, seeds=(20240229, 77, 78)):
        self.g = Dense(t, seeds[0])
        self.a = Dense(t, seeds[1], fock_seed=11)
        self.b = Dense(t, seeds[2], fock_seed=12)
        self.r0, self.excit = r0, excit
        self.x1, self.x2 = self.a.t1, self.a.t2
        self.m1, self.m2 = self.a.n1(), self.a.n2()
        self.m3, self.m4 = self.b.n1(), self.b.n2()
        self.q2 = self.a.e2()

    def six_index(self):
        """(right, left) of cr_eomccsd_t.F:377-419, indexed [p4,p5,p6,h1,h2,h3], from the TCE expressions
        creomccsd_t_n2_mem.F:9,:35,:55,:73 and q3rexpt2.F:7-8."""
        g = self.g
        E = np.einsum
        p9 = Dense._p9
        S, D, M, Et = g.six_index()
        lr0 = abs(self.r0) >= 1e-7
        R = (self.r0 * M if lr0 else 0.0 * M)
        R = R - p9(E("abim,mcjk->abcijk", g.t2, self.m1), 2, 0)          # :9   -P(9) Sum(h12) t(p4 p5 h1 h12) i1(h12 p6 h2 h3)
        R = R + 2.0 * p9(E("aeij,bcke->abcijk", g.t2, self.m2), 0, 2)    # :35  +2 P(9) Sum(p12) t(p4 p12 h1 h2) i1(p5 p6 h3 p12)
        R = R - p9(E("abim,mcjk->abcijk", self.x2, self.m3), 2, 0)       # :55  -P(9) Sum(h12) x(p4 p5 h1 h12) i1(h12 p6 h2 h3)
        R = R - p9(E("aeij,bcke->abcijk", self.x2, self.m4), 0, 2)       # :73  -P(9) Sum(p10) x(p4 p10 h1 h2) i1(p5 p6 h3 p10)
        Lt = (self.r0 * Et if lr0 else 0.0 * Et)
        Lt = Lt + p9(E("abij,ck->abcijk", g.t2, self.x1), 2, 2)          # q3rexpt2.F:7  P(9) t(p4 p5 h1 h2) x(p6 h3)
        Lt = Lt - 2.0 * p9(E("ai,bcjk->abcijk", g.t1, self.q2), 0, 0)    # :8  -2 P(9) t(p4 h1) i1(p5 p6 h2 h3)
        return R, Lt

    def dense_reference(self):
        R, Lt = self.six_index()
        g = self.g
        ep, eh = g.eps[g.P], g.eps[g.H]
        delta = (-ep[:, None, None, None, None, None] - ep[None, :, None, None, None, None] - ep[None, None, :, None, None, None]
                 + eh[None, None, None, :, None, None] + eh[None, None, None, None, :, None] + eh[None, None, None, None, None, :])
        denex = delta + self.excit
        return (float(np.sum(R * R / denex) / 36.0), float(np.sum(Lt * R) / 36.0), float(np.sum(Lt * R / denex) / 36.0),
                float(np.sum(Lt * Lt) / 36.0))

    def _pack(self, d, dense, kind):
        t = d.t
        if kind == "n1":
            h, n = tl.cr_n1_offset(t)
            out = np.zeros(n)
            for key, off in synth._iter_hash(h):
                p4b, h11b, h1b, h2b = tl.decode_cr_n1_key(t, key)
                blk = dense[np.ix_(d._hidx(h11b), d._pidx(p4b), d._hidx(h1b), d._hidx(h2b))].transpose(1, 0, 2, 3)
                out[off:off + blk.size] = blk.ravel()
        elif kind == "n2":
            h, n = tl.cr_n2_offset(t)
            out = np.zeros(n)
            for key, off in synth._iter_hash(h):
                p4b, p5b, h1b, p12b = tl.decode_cr_n2_key(t, key)
                blk = dense[np.ix_(d._pidx(p4b), d._pidx(p5b), d._hidx(h1b), d._pidx(p12b))]
                out[off:off + blk.size] = blk.ravel()
        elif kind == "pphh":
            h, n = tl.t2_offset(t)
            out = np.zeros(n)
            for key, off in synth._iter_hash(h):
                p4b, p5b, h1b, h2b = tl.decode_t2_key(t, key)
                blk = dense[np.ix_(d._pidx(p4b), d._pidx(p5b), d._hidx(h1b), d._hidx(h2b))]
                out[off:off + blk.size] = blk.ravel()
        else:   # "ph": the T1 block structure
            h, n = tl.t1_offset(t)
            out = np.zeros(n)
            for key, off in synth._iter_hash(h):
                p5b, h6b = tl.decode_t1_key(t, key)
                blk = dense[np.ix_(d._pidx(p5b), d._hidx(h6b))]
                out[off:off + blk.size] = blk.ravel()
        return h, out

    def stores(self):
        """(CRStores of the ground-state problem, CREOMStores)"""
        g = self.g
        x1h, x1 = self._pack(g, self.x1, "ph"); x2h, x2 = self._pack(g, self.x2, "pphh")
        m1h, m1 = self._pack(g, self.m1, "n1"); m2h, m2 = self._pack(g, self.m2, "n2")
        m3h, m3 = self._pack(g, self.m3, "n1"); m4h, m4 = self._pack(g, self.m4, "n2")
        q2h, q2 = self._pack(g, self.q2, "pphh")
        return g.stores(), CREOMStores(x1h, x1, x2h, x2, m1h, m1, m2h, m2, m3h, m3, m4h, m4, q2h, q2, self.r0, self.excit)


def cr_energies(num1, num2, den1, den2, den0):
    """cr_ccsd_t.F:260-263."""
    return num1 / (1.0 + den1 + den0), num2 / (1.0 + den2 + den0)
