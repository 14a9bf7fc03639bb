"""Real amplitudes for the reference's own QA case -- TEST INFRASTRUCTURE ONLY.

QA/tests/tce_ccsd_t_h2o (tce_ccsd_t_h2o.nw: H2O, cc-pVDZ spherical, RHF, no frozen core, CCSD(T)) is the one golden
vector set the reference ships for this path: tce_ccsd_t_h2o.out:390 (SCF), :734 (CCSD), :743 / :746 (the
[T] and (T) corrections).  Nothing in the image can produce the converged CCSD amplitudes and MO integrals those numbers
need (no Fortran, no quantum-chemistry package), so this module computes them from first principles in numpy:
McMurchie-Davidson integrals over the cc-pVDZ basis (data: src/basis/libraries/cc-pvdz, "O_cc-pVDZ" / "H_cc-pVDZ"),
RHF, spin-orbital CCSD.  Each stage is checked against the QA output's own intermediate energies (SCF total energy,
CCSD correlation energy) before the amplitudes are handed to the oracle's (T); the oracle's E[T] / E(T) on them is then
compared with the golden corrections (tests/test_qa_h2o.py).  `python -m oracle.h2o_ccsd` regenerates
tests/golden/h2o_ccpvdz_ccsd.npz.  Two more QA cases are generated the same way: ozone (tce_ozone_2eorb / tce_ccsd_t_xmem;
frozen core, `2eorb`; too large to commit, gated) and glycine / STO-3G (tce_lr_ccsd_t: the LR-CCSD(T) energies, which pin
the CR-CCSD(T) tiles; `python -m oracle.h2o_ccsd glycine` -> tests/golden/glycine_sto3g_ccsd.npz, tests/test_qa_lr.py).

Energies are invariant under any change of basis within the same span, so no care is taken to normalise the basis
functions: a contracted function is sum_k c_k a_k^((2l+3)/4) x^i y^j z^k exp(-a_k r^2) (the library's coefficients refer
to normalised primitives: only the exponent dependence of the normalisation matters), d shells are the five real solid
harmonics 2zz-xx-yy, xx-yy, xy, xz, yz."""
from __future__ import annotations
import itertools
import os
import numpy as np
from scipy import special

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(os.path.dirname(HERE), "tests", "golden", "h2o_ccpvdz_ccsd.npz")

# tce_ccsd_t_h2o.nw:17-21 (bohr)
GEOM = [("O", 8.0, (0.0, 0.0, 0.22138519)), ("H", 1.0, (0.0, -1.43013023, -0.88554075)), ("H", 1.0, (0.0, 1.43013023, -0.88554075))]
# src/basis/libraries/cc-pvdz: basis "O_cc-pVDZ" SPHERICAL / "H_cc-pVDZ" SPHERICAL.  (l, exponents, contraction columns)
BASIS = {
    "O": [(0, [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013],
           [[0.00071, 0.00547, 0.027837, 0.1048, 0.283062, 0.448719, 0.270952, 0.015458],
            [-0.00016, -0.001263, -0.006267, -0.025716, -0.070924, -0.165411, -0.116955, 0.557368]]),
          (0, [0.3023], [[1.0]]),
          (1, [17.7, 3.854, 1.046], [[0.043018, 0.228913, 0.508728]]),
          (1, [0.2753], [[1.0]]),
          (2, [1.185], [[1.0]])],
    "H": [(0, [13.01, 1.962, 0.4446], [[0.019685, 0.137977, 0.478148]]),
          (0, [0.122], [[1.0]]),
          (1, [0.727], [[1.0]])],
}
# tce_ccsd_t_h2o.out
QA = dict(scf=-76.026807857236, ccsd_corr=-0.213269954065481, t_bracket=-0.003139909173705, t_paren=-0.003054718622142)

# QA/tests/tce_ozone_2eorb (tce_ozone_2eorb.nw: O3, the inline [5s3p2d] basis, RHF, `freeze atomic`, `2eorb`, CCSD(T)) and
# QA/tests/tce_ccsd_t_xmem (same molecule and basis, symmetry c1, the sliced (T) of ccsd_t_6dts.F under tce:xmem)
OZONE_GEOM = [("O", 8.0, (0.0, 0.0, 0.0)), ("O", 8.0, (0.0, -2.0473224350, -1.2595211660)), ("O", 8.0, (0.0, 2.0473224350, -1.2595211660))]
OZONE_BASIS = {
    "O": [(0, [10662.285, 1599.7097, 364.72526, 103.65179, 33.905805], [[0.000799, 0.006153, 0.031157, 0.115596, 0.301552]]),
          (0, [12.287469, 4.756805], [[0.44487, 0.243172]]),
          (0, [1.004271], [[1.0]]), (0, [0.300686], [[1.0]]), (0, [0.09003], [[1.0]]),
          (1, [34.856463, 7.843131, 2.306249, 0.723164], [[0.015648, 0.098197, 0.307768, 0.49247]]),
          (1, [0.214882], [[1.0]]), (1, [0.06385], [[1.0]]),
          (2, [2.3062, 0.7232], [[0.2027, 0.5791]]),
          (2, [0.2149, 0.0639], [[0.78545, 0.53387]])],
}
# tce_ozone_2eorb.out:396,:898,:908,:911 ; tce_ccsd_t_xmem.out:378,:865,:876,:879
QA_OZONE = dict(scf=-224.327430429177, ccsd_corr=-0.631946819284344, t_bracket=-0.039379872138382, t_paren=-0.036050224214312,
                xmem=dict(scf=-224.327430428908, ccsd_corr=-0.631946818916447, t_bracket=-0.039379871636142, t_paren=-0.036050224479361))
FIXTURE_OZONE = os.path.join(os.path.dirname(HERE), "tests", "golden", "ozone_ccsd.npz")

# QA/tests/tce_lr_ccsd_t (tce_lr_ccsd_t.nw: glycine, STO-3G, RHF, `freeze atomic`, tilesize 10, lr-ccsd(t)): the one QA case
# whose golden numbers depend on the DRESSED moment intermediates of cr_ccsd_t_N and on the t1 (x) t2 tile of cr_ccsd_t_E.
GLYCINE_GEOM = [("O", 8.0, (-2.8770919486, 1.5073755650, 0.3989960497)), ("C", 6.0, (-0.9993929716, 0.2223265108, -0.0939400216)),
                ("C", 6.0, (1.6330980507, 1.1263991128, -0.7236778647)), ("O", 8.0, (-1.3167079358, -2.3304840070, -0.1955378962)),
                ("N", 7.0, (3.5887721300, -0.1900460352, 0.6355723246)), ("H", 1.0, (1.7384347574, 3.1922914768, -0.2011420479)),
                ("H", 1.0, (1.8051078216, 0.9725472539, -2.8503867814)), ("H", 1.0, (3.3674278149, -2.0653924379, 0.5211399625)),
                ("H", 1.0, (5.2887327108, 0.3011058554, -0.0285088728)), ("H", 1.0, (-3.0501350657, -2.7557071585, 0.2342441831))]
_S3, _SP_S, _SP_P = [0.15432897, 0.53532814, 0.44463454], [-0.09996723, 0.39951283, 0.70011547], [0.15591627, 0.60768372, 0.39195739]


def _sto3g(core, valence):
    return [(0, core, [_S3]), (0, valence, [_SP_S]), (1, valence, [_SP_P])]


# src/basis/libraries/sto-3g: "H_STO-3G", "C_STO-3G", "N_STO-3G", "O_STO-3G"
STO3G = {"H": [(0, [3.42525091, 0.62391373, 0.16885540], [_S3])],
         "C": _sto3g([71.6168370, 13.0450960, 3.5305122], [2.9412494, 0.6834831, 0.2222899]),
         "N": _sto3g([99.1061690, 18.0523120, 4.8856602], [3.7804559, 0.8784966, 0.2857144]),
         "O": _sto3g([130.7093200, 23.8088610, 6.4436083], [5.0331513, 1.1695961, 0.3803890])}
# tce_lr_ccsd_t.out:440,:443,:916,:924-939 (correlation energies: CCSD + the LR correction)
QA_GLYCINE = dict(scf=-279.104932121548, enuc=178.409233643692, ccsd_corr=-0.299493468934347,
                  lr=dict(IA=-0.307053292111625, IB=-0.305197418406767, IIA=-0.307599613508071,
                          IIB=-0.305743739803214, IIIA=-0.308248545881664, IIIB=-0.306276971454144))
FIXTURE_GLYCINE = os.path.join(os.path.dirname(HERE), "tests", "golden", "glycine_sto3g_ccsd.npz")

# The H2O / DZ full-CI benchmark (R_e = 1.84345 bohr, HOH = 110.565 deg, both bonds stretched to 1.5 R_e and 2 R_e; Dunning
# [4s2p/2s] basis = src/basis/libraries/dz_dunning "O_DZ (Dunning)" / "H_DZ (Dunning)"; all electrons correlated).  Published
# numbers: RHF and full-CI energies of Olsen, Jorgensen, Koch, Balkova, Bartlett, J. Chem. Phys. 104, 8007 (1996); errors of
# CCSD, CCSD(T) and CR-CCSD(T) relative to full CI (millihartree) from Kowalski, Piecuch, J. Chem. Phys. 113, 18 (2000),
# the paper the reference's manual cites for `cr-ccsd(t)` (doc/user/tce.tex).  The papers are not available in this
# environment: the values were entered from the published tables as remembered, and the check below is their own
# corroboration -- three 8-digit SCF energies and nine 3-decimal errors reproduced to the last digit (tests/test_lit_h2o_dz.py).
DZ_BASIS = {"H": [(0, [19.2406, 2.8992, 0.6534], [[0.032828, 0.231208, 0.817238]]), (0, [0.1776], [[1.0]])],
            "O": [(0, [7816.54, 1175.82, 273.188, 81.1696, 27.1836, 3.4136], [[0.002031, 0.015436, 0.073771, 0.247606, 0.611832, 0.241205]]),
                  (0, [9.5322], [[1.0]]), (0, [0.9398], [[1.0]]), (0, [0.2846], [[1.0]]),
                  (1, [35.1832, 7.904, 2.3051, 0.7171], [[0.01958, 0.124189, 0.394727, 0.627375]]), (1, [0.2137], [[1.0]])]}
H2O_DZ_LIT = {1.0: dict(scf=-76.009838, fci=-76.157866, ccsd=1.790, ccsd_t=0.574, cr_ccsd_t=0.738),
              1.5: dict(scf=-75.803529, fci=-76.014521, ccsd=5.590, ccsd_t=1.465, cr_ccsd_t=2.534),
              2.0: dict(scf=-75.595180, fci=-75.905247, ccsd=9.333, ccsd_t=-7.699, cr_ccsd_t=1.830)}


# The HF / DZ full-CI benchmark of the same paper (R_e = 1.7328 bohr, stretched to 2 R_e and 3 R_e; "F_DZ (Dunning)"): full-CI
# energies and the errors of CCSD, CCSD(T), CR-CCSD(T) in millihartree.  At 3 R_e CCSD(T) is 24.5 millihartree below full CI
# and den0 = 0.98.
DZ_BASIS["F"] = [(0, [9994.79, 1506.03, 350.269, 104.053, 34.8432, 4.3688], [[0.002017, 0.015295, 0.07311, 0.24642, 0.612593, 0.242489]]),
                 (0, [12.2164], [[1.0]]), (0, [1.2078], [[1.0]]), (0, [0.3634], [[1.0]]),
                 (1, [44.3555, 10.082, 2.9959, 0.9383], [[0.020868, 0.130092, 0.396219, 0.620368]]), (1, [0.2733], [[1.0]])]
HF_DZ_LIT = {1.0: dict(fci=-100.160300, ccsd=1.634, ccsd_t=0.325, cr_ccsd_t=0.500),
             2.0: dict(fci=-100.021733, ccsd=6.047, ccsd_t=0.038, cr_ccsd_t=2.031),
             3.0: dict(fci=-99.985281, ccsd=11.596, ccsd_t=-24.480, cr_ccsd_t=2.100)}


def h2o_dz_geometry(stretch=1.0, re=1.84345, angle=110.565):
    R, th = re * stretch, np.deg2rad(angle) / 2
    return [("O", 8.0, (0.0, 0.0, 0.0)), ("H", 1.0, (0.0, R * np.sin(th), R * np.cos(th))), ("H", 1.0, (0.0, -R * np.sin(th), R * np.cos(th)))]


CART = {0: [(0, 0, 0)], 1: [(1, 0, 0), (0, 1, 0), (0, 0, 1)],
        2: [(2, 0, 0), (0, 2, 0), (0, 0, 2), (1, 1, 0), (1, 0, 1), (0, 1, 1)]}
# real solid harmonics of l = 2 over (xx, yy, zz, xy, xz, yz); unnormalised (see the module docstring)
SPH_D = np.array([[-1.0, -1.0, 2.0, 0, 0, 0], [1.0, -1.0, 0, 0, 0, 0], [0, 0, 0, 1.0, 0, 0], [0, 0, 0, 0, 1.0, 0], [0, 0, 0, 0, 0, 1.0]])


def check_basis_against_reference(path="/root/reference/src/basis/libraries/cc-pvdz", name="cc-pVDZ", basis=None, elements=("O", "H")):
    """The numbers above are the library's (only where the reference tree is mounted).  An SP shell of the library is two
    entries here (s and p over the same exponents), so exponents are compared as sets and coefficients as multisets."""
    if not os.path.exists(path):
        return None
    basis = BASIS if basis is None else basis
    txt = open(path).read()
    for el in elements:
        blk = txt[txt.index(f'basis "{el}_{name}"'):]
        blk = blk[:blk.index("\nend")]
        rows = [[float(x) for x in line.split()] for line in blk.split("\n")[1:] if line.split() and (line.split()[0][0].isdigit() or line.split()[0][0] == "-")]
        ref_exp = sorted({r[0] for r in rows})
        ref_cof = sorted(c for r in rows for c in r[1:])
        my_exp = sorted({a for l, exps, cols in basis[el] for a in exps})
        my_cof = sorted(c[k] for l, exps, cols in basis[el] for k in range(len(exps)) for c in cols)
        assert np.allclose(ref_exp, my_exp, rtol=0, atol=1e-12), el
        assert np.allclose(ref_cof, my_cof, rtol=0, atol=1e-12), el
    return True


# ------------------------------------------------------------------------------------------------------------------
# primitive Cartesian functions and the contraction to the 24 spherical basis functions
# ------------------------------------------------------------------------------------------------------------------
def build_basis(geom=None, basis=None):
    geom = GEOM if geom is None else geom
    basis = BASIS if basis is None else basis
    prim = []      # (center, exponent, (i,j,k))
    rows = []      # contraction: list of (basis function index, prim index, coefficient)
    nbf = 0
    for el, z, R in geom:
        for l, exps, cols in basis[el]:
            first = len(prim)
            for a in exps:
                for ijk in CART[l]:
                    prim.append((np.array(R), a, ijk))
            ncart = len(CART[l])
            for col in cols:
                comb = np.eye(ncart) if l < 2 else SPH_D
                for f in range(comb.shape[0]):
                    for k, a in enumerate(exps):
                        w = col[k] * a ** ((2 * l + 3) / 4.0)
                        for cidx in range(ncart):
                            if comb[f, cidx] != 0.0:
                                rows.append((nbf, first + k * ncart + cidx, w * comb[f, cidx]))
                    nbf += 1
    C = np.zeros((len(prim), nbf))
    for bf, p, w in rows:
        C[p, bf] += w
    return prim, C


def _e1d(i, j, a, b, Q):
    """Hermite expansion coefficients E_t^{ij}, t = 0..i+j, of x_A^i x_B^j exp(-a x_A^2 - b x_B^2) (Q = A - B)."""
    p = a + b
    q = a * b / p
    memo = {}

    def E(i, j, t):
        if t < 0 or t > i + j:
            return 0.0
        key = (i, j, t)
        if key in memo:
            return memo[key]
        if i == 0 and j == 0:
            v = np.exp(-q * Q * Q)
        elif j == 0:
            v = E(i - 1, j, t - 1) / (2 * p) - (q * Q / a) * E(i - 1, j, t) + (t + 1) * E(i - 1, j, t + 1)
        else:
            v = E(i, j - 1, t - 1) / (2 * p) + (q * Q / b) * E(i, j - 1, t) + (t + 1) * E(i, j - 1, t + 1)
        memo[key] = v
        return v
    return [E(i, j, t) for t in range(i + j + 1)]


HERM = [h for L in range(5) for h in itertools.product(range(L + 1), repeat=3) if sum(h) == L]      # t+u+v <= 4: 35 terms
HIDX = {h: n for n, h in enumerate(HERM)}


def boys(nmax, T):
    """F_n(T), n = 0..nmax, for an array T >= 0: downward recursion from F_nmax (incomplete gamma function; series near 0)."""
    T = np.asarray(T, dtype=np.float64)
    out = np.zeros((nmax + 1,) + T.shape)
    small = T < 1e-6
    Ts = np.where(small, 1.0, T)
    n = nmax
    top = special.gammainc(n + 0.5, Ts) * special.gamma(n + 0.5) / (2.0 * Ts ** (n + 0.5))
    top = np.where(small, 1.0 / (2 * n + 1) - T / (2 * n + 3) + T * T / (2 * (2 * n + 5)), top)
    out[n] = top
    eT = np.exp(-T)
    for m in range(nmax - 1, -1, -1):
        out[m] = (2.0 * T * out[m + 1] + eT) / (2 * m + 1)
    return out


def hermite_R(Lmax, alpha, X, Y, Z):
    """R_{tuv} = R^0_{tuv}(alpha, (X,Y,Z)) for t+u+v <= Lmax, vectorised over the leading axis; returns dict (t,u,v) -> array."""
    T = alpha * (X * X + Y * Y + Z * Z)
    F = boys(Lmax, T)
    R = {}
    for n in range(Lmax + 1):
        R[(0, 0, 0, n)] = (-2.0 * alpha) ** n * F[n]
    for L in range(1, Lmax + 1):
        for t, u, v in itertools.product(range(L + 1), repeat=3):
            if t + u + v != L:
                continue
            for n in range(Lmax - L + 1):
                if t > 0:
                    val = X * R[(t - 1, u, v, n + 1)]
                    if t > 1:
                        val = val + (t - 1) * R[(t - 2, u, v, n + 1)]
                elif u > 0:
                    val = Y * R[(t, u - 1, v, n + 1)]
                    if u > 1:
                        val = val + (u - 1) * R[(t, u - 2, v, n + 1)]
                else:
                    val = Z * R[(t, u, v - 1, n + 1)]
                    if v > 1:
                        val = val + (v - 1) * R[(t, u, v - 2, n + 1)]
                R[(t, u, v, n)] = val
    return {k[:3]: v for k, v in R.items() if k[3] == 0}


def integrals(geom=None, basis=None, screen=1e-16, chunk=16000, verbose=False):
    """(S, T, V, eri, Enuc) over the contracted spherical functions; eri[p,q,r,s] = (pq|rs).  Primitive pairs whose
    Hermite coefficients are all below `screen` (tight functions on different centres) are dropped."""
    geom = GEOM if geom is None else geom
    prim, C = build_basis(geom, basis)
    n = len(prim)
    nbf = C.shape[1]
    pairs = [(i, j) for i in range(n) for j in range(i + 1)]
    npair = len(pairs)
    P = np.zeros((npair, 3)); pp = np.zeros(npair); EH = np.zeros((npair, len(HERM)))
    S = np.zeros((n, n)); Tk = np.zeros((n, n))
    nuc = [(z, np.array(R)) for _, z, R in geom]
    for k, (i, j) in enumerate(pairs):
        A, a, la = prim[i]; B, b, lb = prim[j]
        p = a + b
        Pc = (a * A + b * B) / p
        e = [_e1d(la[d], lb[d], a, b, A[d] - B[d]) for d in range(3)]
        for (t, u, v), idx in HIDX.items():
            if t <= la[0] + lb[0] and u <= la[1] + lb[1] and v <= la[2] + lb[2]:
                EH[k, idx] = e[0][t] * e[1][u] * e[2][v]
        P[k] = Pc; pp[k] = p
        # overlap and kinetic energy from one-dimensional overlaps s_d(i,j) = E_0^{ij} sqrt(pi/p)
        def s1(d, ii, jj):
            if ii < 0 or jj < 0:
                return 0.0
            return _e1d(ii, jj, a, b, A[d] - B[d])[0] * np.sqrt(np.pi / p)
        s = [s1(d, la[d], lb[d]) for d in range(3)]
        t1 = [-2 * b * b * s1(d, la[d], lb[d] + 2) + b * (2 * lb[d] + 1) * s[d] - 0.5 * lb[d] * (lb[d] - 1) * s1(d, la[d], lb[d] - 2)
              for d in range(3)]
        S[i, j] = S[j, i] = s[0] * s[1] * s[2]
        Tk[i, j] = Tk[j, i] = t1[0] * s[1] * s[2] + s[0] * t1[1] * s[2] + s[0] * s[1] * t1[2]
    # nuclear attraction, all pairs at once per nucleus: V_ij = -Z (2 pi / p) sum_tuv E_tuv R_tuv(p, P - C)
    vpair = np.zeros(npair)
    for z, Rc in nuc:
        d = P - Rc[None, :]
        R = hermite_R(4, pp, d[:, 0], d[:, 1], d[:, 2])
        vpair += -z * (2 * np.pi / pp) * sum(EH[:, idx] * R[h] for h, idx in HIDX.items())
    V = np.zeros((n, n))
    for k, (i, j) in enumerate(pairs):
        V[i, j] = V[j, i] = vpair[k]
    # contraction of primitive pairs to contracted pairs: W[k, (a>=b)] so that (ab| = sum_k W[k,ab] (k|
    ia, ib = np.tril_indices(nbf)
    W = np.zeros((npair, len(ia)))
    for k, (i, j) in enumerate(pairs):
        w = C[i][ia] * C[j][ib]
        if i != j:
            w = w + C[j][ia] * C[i][ib]
        W[k] = w
    keep = np.where((np.max(np.abs(EH), axis=1) > screen) & (np.max(np.abs(W), axis=1) > 0))[0]
    if verbose:
        print(f"primitive functions {n}, pairs {npair}, kept {len(keep)}", flush=True)
    # sort the kept pairs by total angular momentum: a pair of L needs only the (L+1)(L+2)(L+3)/6 leading Hermite terms
    Lpair = np.array([sum(prim[i][2]) + sum(prim[j][2]) for (i, j) in pairs])[keep]
    order = np.argsort(Lpair, kind="stable")
    keep = keep[order]; Lpair = Lpair[order]
    EHk, Pk, ppk, Wk = EH[keep], P[keep], pp[keep], W[keep]
    nk = len(keep)
    nherm = [(L + 1) * (L + 2) * (L + 3) // 6 for L in range(5)]
    groups = [np.where(Lpair == L)[0] for L in range(5)]
    sign = np.array([(-1.0) ** sum(h) for h in HERM])
    EHs = EHk * sign[None, :]
    npc = len(ia)
    half = np.zeros((nk, npc))                                            # half[a, (cd)] = sum_b (a|b) W[b, cd]
    tables = {}

    def table(LA, LB):
        if (LA, LB) not in tables:
            hs, gs = HERM[:nherm[LA]], HERM[:nherm[LB]]
            keys = sorted({tuple(np.add(h, g)) for h in hs for g in gs})
            kpos = {kk: m for m, kk in enumerate(keys)}
            tables[(LA, LB)] = (keys, np.array([[kpos[tuple(np.add(h, g))] for g in gs] for h in hs]))
        return tables[(LA, LB)]

    def block(ra, rb, LA, LB):
        """(a|b) for every a in ra, b in rb (index arrays into the kept pairs): matrix [len(ra), len(rb)]"""
        keys, gather = table(LA, LB)
        A = np.repeat(ra, len(rb)); B = np.tile(rb, len(ra))
        p = ppk[A]; q = ppk[B]
        alpha = p * q / (p + q)
        d = Pk[A] - Pk[B]
        R = hermite_R(LA + LB, alpha, d[:, 0], d[:, 1], d[:, 2])
        Rm = np.stack([R[kk] for kk in keys], axis=1)
        val = np.einsum("bh,bhg,bg->b", EHk[A][:, :nherm[LA]], Rm[:, gather], EHs[B][:, :nherm[LB]], optimize=True)
        val *= 2.0 * np.pi ** 2.5 / (p * q * np.sqrt(p + q))
        return val.reshape(len(ra), len(rb))

    done = 0
    for LA in range(5):
        for LB in range(LA + 1):
            ga, gb = groups[LA], groups[LB]
            if len(ga) == 0 or len(gb) == 0:
                continue
            per = max(1, int(chunk * 1500 / (len(gb) * nherm[LA] * nherm[LB])))   # ~ chunk * 1500 gathered R values per batch
            for a0 in range(0, len(ga), per):
                ra = ga[a0:a0 + per]
                if LA == LB:
                    rb = gb[:a0 + len(ra)]                                # lower triangle of the diagonal class block
                else:
                    rb = gb
                val = block(ra, rb, LA, LB)
                if LA == LB:
                    # rows ra against columns rb (which end with ra itself): count the ra x ra part once
                    nlow = a0
                    half[ra] += val @ Wk[rb]
                    if nlow:
                        half[rb[:nlow]] += val[:, :nlow].T @ Wk[ra]
                else:
                    half[ra] += val @ Wk[rb]
                    half[rb] += val.T @ Wk[ra]
                done += val.size
            if verbose:
                print(f"  eri class ({LA},{LB}): {len(ga)} x {len(gb)} pairs", flush=True)
    full = Wk.T @ half                                                    # [(ab), (cd)]
    pidx = np.zeros((nbf, nbf), dtype=np.int64)
    pidx[ia, ib] = np.arange(npc); pidx[ib, ia] = np.arange(npc)
    eri = full[pidx.reshape(-1)][:, pidx.reshape(-1)].reshape(nbf, nbf, nbf, nbf)
    eri = 0.5 * (eri + eri.transpose(2, 3, 0, 1))
    Sc = C.T @ S @ C; Tc = C.T @ Tk @ C; Vc = C.T @ V @ C
    enuc = sum(nuc[i][0] * nuc[j][0] / np.linalg.norm(nuc[i][1] - nuc[j][1]) for i in range(len(nuc)) for j in range(i))
    return Sc, Tc, Vc, eri, float(enuc)


def rhf(S, T, V, eri, enuc, nocc=5, tol=1e-12, maxit=200):
    """Restricted Hartree-Fock with DIIS; returns (energy, orbital energies, MO coefficients)."""
    from scipy import linalg
    H = T + V
    e, Cm = linalg.eigh(H, S)
    D = Cm[:, :nocc] @ Cm[:, :nocc].T
    fs, es = [], []
    E = 0.0
    for it in range(maxit):
        J = np.einsum("pqrs,rs->pq", eri, D)
        K = np.einsum("prqs,rs->pq", eri, D)
        F = H + 2 * J - K
        Enew = float(np.sum(D * (H + F))) + enuc
        err = F @ D @ S - S @ D @ F
        fs.append(F); es.append(err)
        fs, es = fs[-8:], es[-8:]
        if len(fs) > 1:
            m = len(fs)
            Bm = -np.ones((m + 1, m + 1)); Bm[m, m] = 0.0
            for i in range(m):
                for j in range(m):
                    Bm[i, j] = np.sum(es[i] * es[j])
            rhs = np.zeros(m + 1); rhs[m] = -1.0
            c = np.linalg.solve(Bm, rhs)[:m]
            F = sum(ci * fi for ci, fi in zip(c, fs))
        e, Cm = linalg.eigh(F, S)
        D = Cm[:, :nocc] @ Cm[:, :nocc].T
        if abs(Enew - E) < tol and np.max(np.abs(err)) < 1e-9:
            E = Enew
            break
        E = Enew
    return E, e, Cm


def ccsd(eps, eri_mo, nocc=5, tol=1e-11, maxit=200, rms_tol=1e-10):
    """Spin-orbital CCSD (Stanton, Gauss, Watts, Bartlett, J. Chem. Phys. 94, 4334 (1991)) for a canonical closed-shell RHF
    reference; returns (correlation energy, t1s[a,i], t2s[a,b,i,j]) -- the spatial amplitudes t_i^a, t_{ij}^{ab} (alpha-beta)."""
    n = len(eps)
    nso = 2 * n
    # spin-orbital index = 2*spatial + spin
    sp = np.arange(nso) // 2
    spin = np.arange(nso) % 2
    g = eri_mo[np.ix_(sp, sp, sp, sp)]                                     # (pq|rs) over spin orbitals, spatial part
    dl = (spin[:, None] == spin[None, :]).astype(float)
    # <pq||rs> = (pr|qs) d(p,r) d(q,s) - (ps|qr) d(p,s) d(q,r)
    v = np.einsum("prqs,pr,qs->pqrs", g, dl, dl) - np.einsum("psqr,ps,qr->pqrs", g, dl, dl)
    e = eps[sp]
    no = 2 * nocc
    o, vv = slice(0, no), slice(no, nso)
    eo, ev = e[o], e[vv]
    D1 = eo[:, None] - ev[None, :]
    D2 = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, :]
    t1 = np.zeros((no, nso - no))
    t2 = v[o, o, vv, vv] / D2

    def E(*a):
        return np.einsum(*a, optimize=True)

    def energy(t1, t2):
        return 0.25 * E("ijab,ijab->", v[o, o, vv, vv], t2) + 0.5 * E("ijab,ia,jb->", v[o, o, vv, vv], t1, t1)
    t1s_l, t2s_l, err_l = [], [], []
    Ecc = energy(t1, t2)
    for it in range(maxit):
        tau_t = t2 + 0.5 * (E("ia,jb->ijab", t1, t1) - E("ib,ja->ijab", t1, t1))
        tau = t2 + E("ia,jb->ijab", t1, t1) - E("ib,ja->ijab", t1, t1)
        Fae = (-0.5 * E("me,ma->ae", np.zeros_like(t1), t1) + E("mf,mafe->ae", t1, v[o, vv, vv, vv])
               - 0.5 * E("mnaf,mnef->ae", tau_t, v[o, o, vv, vv]))
        Fmi = (E("ne,mnie->mi", t1, v[o, o, o, vv]) + 0.5 * E("inef,mnef->mi", tau_t, v[o, o, vv, vv]))
        Fme = E("nf,mnef->me", t1, v[o, o, vv, vv])
        Wmnij = (v[o, o, o, o] + E("je,mnie->mnij", t1, v[o, o, o, vv]) - E("ie,mnje->mnij", t1, v[o, o, o, vv])
                 + 0.25 * E("ijef,mnef->mnij", tau, v[o, o, vv, vv]))
        Wabef = (v[vv, vv, vv, vv] - E("mb,amef->abef", t1, v[vv, o, vv, vv]) + E("ma,bmef->abef", t1, v[vv, o, vv, vv])
                 + 0.25 * E("mnab,mnef->abef", tau, v[o, o, vv, vv]))
        Wmbej = (v[o, vv, vv, o] + E("jf,mbef->mbej", t1, v[o, vv, vv, vv]) - E("nb,mnej->mbej", t1, v[o, o, vv, o])
                 - E("jnfb,mnef->mbej", 0.5 * t2 + E("jf,nb->jnfb", t1, t1), v[o, o, vv, vv]))
        # T1 (canonical orbitals: f_ia = 0, f diagonal)
        r1 = (E("ie,ae->ia", t1, Fae) - E("ma,mi->ia", t1, Fmi) + E("imae,me->ia", t2, Fme)
              - E("nf,naif->ia", t1, v[o, vv, o, vv]) - 0.5 * E("imef,maef->ia", t2, v[o, vv, vv, vv])
              - 0.5 * E("mnae,nmei->ia", t2, v[o, o, vv, o]))
        # T2
        Fae2 = Fae - 0.5 * E("mb,me->be", t1, Fme)
        Fmi2 = Fmi + 0.5 * E("je,me->mj", t1, Fme)
        r2 = v[o, o, vv, vv].copy()
        x = E("ijae,be->ijab", t2, Fae2); r2 += x - x.transpose(0, 1, 3, 2)
        x = E("imab,mj->ijab", t2, Fmi2); r2 -= x - x.transpose(1, 0, 2, 3)
        r2 += 0.5 * E("mnab,mnij->ijab", tau, Wmnij) + 0.5 * E("ijef,abef->ijab", tau, Wabef)
        x = E("imae,mbej->ijab", t2, Wmbej) - E("ie,ma,mbej->ijab", t1, t1, v[o, vv, vv, o])
        r2 += x - x.transpose(1, 0, 2, 3) - x.transpose(0, 1, 3, 2) + x.transpose(1, 0, 3, 2)
        x = E("ie,abej->ijab", t1, v[vv, vv, vv, o]); r2 += x - x.transpose(1, 0, 2, 3)
        x = E("ma,mbij->ijab", t1, v[o, vv, o, o]); r2 -= x - x.transpose(0, 1, 3, 2)
        t1n = r1 / D1
        t2n = r2 / D2
        # DIIS on the amplitudes
        t1s_l.append(t1n); t2s_l.append(t2n); err_l.append(np.concatenate([(t1n - t1).ravel(), (t2n - t2).ravel()]))
        t1s_l, t2s_l, err_l = t1s_l[-8:], t2s_l[-8:], err_l[-8:]
        if len(err_l) > 1:
            m = len(err_l)
            Bm = -np.ones((m + 1, m + 1)); Bm[m, m] = 0.0
            for i in range(m):
                for j in range(m):
                    Bm[i, j] = err_l[i] @ err_l[j]
            rhs = np.zeros(m + 1); rhs[m] = -1.0
            c = np.linalg.solve(Bm, rhs)[:m]
            t1n = sum(ci * x for ci, x in zip(c, t1s_l)); t2n = sum(ci * x for ci, x in zip(c, t2s_l))
        dE = energy(t1n, t2n) - Ecc
        rms = np.sqrt(err_l[-1] @ err_l[-1])
        t1, t2 = t1n, t2n
        Ecc += dE
        if abs(dE) < tol and rms < rms_tol:
            break
    # spatial amplitudes: alpha = even spin orbitals
    oa = np.arange(0, no, 2); ob = oa + 1
    va = np.arange(0, nso - no, 2); vb = va + 1
    t1s = t1[np.ix_(oa, va)].T.copy()                                       # [a,i]
    t2s = t2[np.ix_(oa, ob, va, vb)].transpose(2, 3, 0, 1).copy()           # t_{i alpha j beta}^{a alpha b beta} -> [a,b,i,j]
    return float(Ecc), t1s, t2s


def mo_irreps(S, Cm, geom=None, basis=None, inplane_code=2):
    """C2v irrep code (0 a1, 1 a2, 2 / 3 the two b irreps; XOR = direct product) of every MO from its parities under
    x -> -x and y -> -y.  The molecule lies in the yz plane; y -> -y swaps the off-axis atoms.  `inplane_code` goes to the
    in-plane antisymmetric irrep: 2 for H2O (the one with six virtual orbitals, b1 in the QA tile table
    tce_ccsd_t_h2o.out:644-659), 3 for ozone (b2 in tce_ozone_2eorb.out)."""
    geom = GEOM if geom is None else geom
    basis = BASIS if basis is None else basis
    labels = []       # per basis function: (atom, shell key, power of x, power of y)
    for ia, (el, z, R) in enumerate(geom):
        for ish, (l, exps, cols) in enumerate(basis[el]):
            comps = {0: [(0, 0)], 1: [(1, 0), (0, 1), (0, 0)], 2: [(0, 0), (0, 0), (1, 1), (1, 0), (0, 1)]}[l]
            for icol in range(len(cols)):
                for ic, (px, py) in enumerate(comps):
                    labels.append((ia, (ish, icol, ic), px, py))
    n = len(labels)
    Px = np.zeros((n, n)); Py = np.zeros((n, n))
    swap = {}         # image of every atom under y -> -y
    for ia, (el, z, R) in enumerate(geom):
        swap[ia] = [ja for ja, (_, _, Q) in enumerate(geom) if abs(Q[0] - R[0]) < 1e-9 and abs(Q[1] + R[1]) < 1e-9 and abs(Q[2] - R[2]) < 1e-9][0]
    for i, (ia, key, px, py) in enumerate(labels):
        Px[i, i] = (-1.0) ** px
        j = [k for k, (ja, kk, _, _) in enumerate(labels) if ja == swap[ia] and kk == key][0]
        Py[j, i] = (-1.0) ** py
    out = []
    for m in range(Cm.shape[1]):
        c = Cm[:, m]
        ex = float(c @ S @ (Px @ c)); ey = float(c @ S @ (Py @ c))
        assert abs(abs(ex) - 1.0) < 1e-6 and abs(abs(ey) - 1.0) < 1e-6, (m, ex, ey)
        sx, sy = ex > 0, ey > 0
        inp, outp = inplane_code, 5 - inplane_code
        out.append(0 if (sx and sy) else 1 if (not sx and not sy) else inp if (sx and not sy) else outp)
    return np.array(out, dtype=np.int64)


def generate_ozone(verbose=True):
    """The ozone case of QA/tests/tce_ozone_2eorb and tce_ccsd_t_xmem: 72 basis functions, 12 occupied orbitals of which the
    three O 1s cores are frozen (`freeze atomic`), 60 virtuals.  About 15 minutes and 15 GB; the result is too large for a
    committed fixture (the <pp||hp> class alone is 8 MB), so tests/test_qa_h2o.py runs it only when NWC_QA_OZONE=1 and
    the log of that run is kept in profiles/."""
    S, T, V, eri, enuc = integrals(OZONE_GEOM, OZONE_BASIS, verbose=verbose)
    escf, eps, Cm = rhf(S, T, V, eri, enuc, nocc=12)
    if verbose:
        print(f"SCF   {escf:.12f}   QA {QA_OZONE['scf']:.12f}   diff {escf - QA_OZONE['scf']:.2e}", flush=True)
    irrep = mo_irreps(S, Cm, OZONE_GEOM, OZONE_BASIS, inplane_code=3)
    nfz = 3
    Ca = Cm[:, nfz:]
    eri_mo = np.einsum("pqrs,pa,qb,rc,sd->abcd", eri, Ca, Ca, Ca, Ca, optimize=True)
    ecc, t1s, t2s = ccsd(eps[nfz:], eri_mo, nocc=12 - nfz)
    if verbose:
        print(f"CCSD  {ecc:.12f}   QA {QA_OZONE['ccsd_corr']:.12f}   diff {ecc - QA_OZONE['ccsd_corr']:.2e}", flush=True)
    return dict(escf=escf, ecc=ecc, eps=eps[nfz:], irrep=irrep[nfz:], t1s=t1s, t2s=t2s, eri_mo=eri_mo, nocc=12 - nfz)


def generate_glycine(verbose=True):
    """The glycine / STO-3G case of QA/tests/tce_lr_ccsd_t: 30 basis functions, 20 occupied orbitals of which the five 1s
    cores are frozen, 10 virtuals, no symmetry.  Seconds."""
    check_basis_against_reference("/root/reference/src/basis/libraries/sto-3g", "STO-3G", STO3G, ("H", "C", "N", "O"))
    S, T, V, eri, enuc = integrals(GLYCINE_GEOM, STO3G, verbose=False)
    escf, eps, Cm = rhf(S, T, V, eri, enuc, nocc=20)
    if verbose:
        print(f"Enuc  {enuc:.12f}   QA {QA_GLYCINE['enuc']:.12f}   diff {enuc - QA_GLYCINE['enuc']:.2e}")
        print(f"SCF   {escf:.12f}   QA {QA_GLYCINE['scf']:.12f}   diff {escf - QA_GLYCINE['scf']:.2e}", flush=True)
    nfz = 5
    Ca = Cm[:, nfz:]
    eri_mo = np.einsum("pqrs,pa,qb,rc,sd->abcd", eri, Ca, Ca, Ca, Ca, optimize=True)
    ecc, t1s, t2s = ccsd(eps[nfz:], eri_mo, nocc=20 - nfz)
    if verbose:
        print(f"CCSD  {ecc:.12f}   QA {QA_GLYCINE['ccsd_corr']:.12f}   diff {ecc - QA_GLYCINE['ccsd_corr']:.2e}", flush=True)
    return dict(escf=escf, ecc=ecc, eps=eps[nfz:], irrep=np.zeros(25, dtype=np.int64), t1s=t1s, t2s=t2s, eri_mo=eri_mo, nocc=20 - nfz)


def _generate_dz(geom):
    S, T, V, eri, enuc = integrals(geom, DZ_BASIS)
    escf, eps, Cm = rhf(S, T, V, eri, enuc, nocc=5)
    eri_mo = np.einsum("pqrs,pa,qb,rc,sd->abcd", eri, Cm, Cm, Cm, Cm, optimize=True)
    ecc, t1s, t2s = ccsd(eps, eri_mo, nocc=5, maxit=800)
    return dict(escf=escf, ecc=ecc, eps=eps, irrep=np.zeros(len(eps), dtype=np.int64), t1s=t1s, t2s=t2s, eri_mo=eri_mo, nocc=5)


def generate_h2o_dz(stretch=1.0):
    """SCF + all-electron CCSD of the H2O / DZ benchmark at R = stretch * R_e (14 basis functions; under a second)."""
    return _generate_dz(h2o_dz_geometry(stretch))


def generate_hf_dz(stretch=1.0, re=1.7328):
    """The same for hydrogen fluoride (12 basis functions)."""
    return _generate_dz([("F", 9.0, (0.0, 0.0, 0.0)), ("H", 1.0, (0.0, 0.0, re * stretch))])


def generate(verbose=True):
    check_basis_against_reference()
    S, T, V, eri, enuc = integrals()
    escf, eps, Cm = rhf(S, T, V, eri, enuc)
    irrep = mo_irreps(S, Cm)
    if verbose:
        print(f"SCF   {escf:.12f}   QA {QA['scf']:.12f}   diff {escf - QA['scf']:.2e}", flush=True)
    eri_mo = np.einsum("pqrs,pa,qb,rc,sd->abcd", eri, Cm, Cm, Cm, Cm, optimize=True)
    ecc, t1s, t2s = ccsd(eps, eri_mo)
    if verbose:
        print(f"CCSD  {ecc:.12f}   QA {QA['ccsd_corr']:.12f}   diff {ecc - QA['ccsd_corr']:.2e}", flush=True)
    return dict(escf=escf, ecc=ecc, eps=eps, irrep=irrep, t1s=t1s, t2s=t2s, eri_mo=eri_mo)


def _pack(eri):
    """(pq|rs) -> the upper triangle of the (p>=q) x (r>=s) pair matrix (8-fold symmetry: 45 150 of 331 776 numbers)"""
    n = eri.shape[0]
    i, j = np.tril_indices(n)
    m = eri[i, j][:, i, j]
    a, b = np.triu_indices(len(i))
    return m[a, b]


def _unpack(packed, n):
    i, j = np.tril_indices(n)
    npair = len(i)
    m = np.zeros((npair, npair))
    a, b = np.triu_indices(npair)
    m[a, b] = packed
    m[b, a] = packed
    pidx = np.zeros((n, n), dtype=np.int64)
    pidx[i, j] = np.arange(npair); pidx[j, i] = np.arange(npair)
    return m[pidx.reshape(-1)][:, pidx.reshape(-1)].reshape(n, n, n, n)


def save(r, path=FIXTURE):
    extra = {"nocc": r["nocc"]} if "nocc" in r else {}
    np.savez_compressed(path, escf=r["escf"], ecc=r["ecc"], eps=r["eps"], irrep=r["irrep"], t1s=r["t1s"], t2s=r["t2s"],
                        eri_packed=_pack(r["eri_mo"]), **extra)


def load(path=FIXTURE):
    """The committed fixture (generated by this module): dict(escf, ecc, eps[24], irrep[24], t1s[19,5], t2s[19,19,5,5],
    eri_mo[24,24,24,24] = (pq|rs) over the canonical RHF orbitals in energy order)."""
    d = np.load(path)
    out = {k: d[k] for k in d.files if k != "eri_packed"}
    out["eri_mo"] = _unpack(d["eri_packed"], len(d["eps"]))
    return out


def qa_stores(r=None, tilesize=20, c2v=True, restricted=True, intorb=False):
    """TCE block stores of the QA case from the real tensors.  c2v: the tiling of the QA run itself (the tile table of
    tce_ccsd_t_h2o.out:644-659: occupied a1 3, b1 1, b2 1; virtual a1 8, a2 2, b1 6, b2 3 per spin); otherwise C1."""
    from nwchem_b200 import synth, tiling as tl
    r = load() if r is None else r
    eps, irr = r["eps"], r["irrep"]
    no = int(r["nocc"]) if "nocc" in r else 5
    nv = len(eps) - no
    if c2v:
        occ = [int(np.sum(irr[:no] == g)) for g in range(4)]
        virt = [int(np.sum(irr[no:] == g)) for g in range(4)]
        oo = np.concatenate([np.where(irr[:no] == g)[0] for g in range(4)])          # occupied, grouped by irrep
        vo = np.concatenate([np.where(irr[no:] == g)[0] for g in range(4)])          # virtual, grouped by irrep
    else:
        occ, virt = [no], [nv]
        oo, vo = np.arange(no), np.arange(nv)
    perm = np.concatenate([oo, no + vo])
    t = tl.make_tiling(occ, virt, tilesize, restricted, evl=(eps[:no][oo], eps[no:][vo]))
    t1s = r["t1s"][np.ix_(vo, oo)]
    t2s = r["t2s"][np.ix_(vo, vo, oo, oo)]
    eri = r["eri_mo"][np.ix_(perm, perm, perm, perm)]
    return synth.physical(t, intorb=intorb, dense=(no, nv, t1s, t2s, eri))


if __name__ == "__main__":
    import sys
    if "glycine" in sys.argv[1:]:
        save(generate_glycine(), FIXTURE_GLYCINE)
        print("wrote", FIXTURE_GLYCINE, os.path.getsize(FIXTURE_GLYCINE), "bytes")
    else:
        r = generate()
        save(r)
        print("wrote", FIXTURE, os.path.getsize(FIXTURE), "bytes")
