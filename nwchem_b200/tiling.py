"""Tile tables and TCE block-store offset tables for the (T) path (host-side integer state).

Mirrors, for the state the (T) driver reads:
  * tce_tile            src/tce/tce_tile.F:330-357 (tile split), :1279-1312 (k_spin/k_sym/k_range/k_offset/k_alpha)
  * tce_t1_offset_new   src/tce/tce_t1_offset_new.F:40-53
  * tce_t2_offset_new   src/tce/tce_t2_offset_new.F:45-68
  * tce_mo2e_offset     src/tce/tce_mo2e_offset.F:45-67 (restricted here to the three V2 classes (T) reads)
All tile ids are 1-based as in the reference; arrays are indexed [tile-1].
"""
from __future__ import annotations
import dataclasses
import numpy as np

I64 = np.int64


def tile_group(n: int, isize: int) -> list[int]:
    """tce_tile.F:330-357: split n orbitals of one (spin, irrep) group into ceil(n/isize) near-equal tiles."""
    if n <= 0:
        return []
    nblocks = n // isize
    if n > isize * nblocks:
        nblocks += 1
    out, done = [], 0
    for k in range(1, nblocks + 1):
        r = k * n // nblocks - done
        done += r
        out.append(r)
    return out


@dataclasses.dataclass
class Tiling:
    noab: int
    nvab: int
    restricted: bool
    spin: np.ndarray      # 1 alpha, 2 beta
    sym: np.ndarray       # irrep bit code
    range: np.ndarray
    offset: np.ndarray    # into evl_sorted
    alpha: np.ndarray     # 1-based id of the alpha twin (k_alpha)
    evl_sorted: np.ndarray
    # bookkeeping for synthetic generators: spatial orbital ids of each tile's members
    members: list

    @property
    def ntiles(self) -> int:
        return self.noab + self.nvab

    def r(self, b: int) -> int:
        return int(self.range[b - 1])


def make_tiling(occ_by_irrep: list[int], virt_by_irrep: list[int], tilesize: int, restricted: bool = True,
                seed: int = 20240229, evl: tuple | None = None) -> Tiling:
    """RHF-style tiling: tile order h-alpha (by irrep), h-beta, p-alpha, p-beta (tce_tile.F:300-460).

    Spatial orbital ids: occupied 0..no-1 grouped by irrep, virtual no..no+nv-1 grouped by irrep.
    `evl` = optional (eps_occ, eps_virt) spatial arrays in that order; default seeded synthetic.
    """
    no, nv = sum(occ_by_irrep), sum(virt_by_irrep)
    rng = np.random.default_rng(seed)
    if evl is None:
        eo = np.sort(rng.uniform(-2.0, -0.4, no))
        ev = np.sort(rng.uniform(0.1, 3.0, nv))
    else:
        eo, ev = np.asarray(evl[0], float), np.asarray(evl[1], float)
    spin, sym, rng_, members = [], [], [], []

    def add_group(counts, base, sp):
        start = base
        n_tiles = 0
        for irrep, n in enumerate(counts):
            done = 0
            for r in tile_group(n, tilesize):
                spin.append(sp); sym.append(irrep); rng_.append(r)
                members.append(np.arange(start + done, start + done + r))
                done += r
                n_tiles += 1
            start += n
        return n_tiles

    noa = add_group(occ_by_irrep, 0, 1)
    nob = add_group(occ_by_irrep, 0, 2)
    nva = add_group(virt_by_irrep, no, 1)
    nvb = add_group(virt_by_irrep, no, 2)
    ntile = noa + nob + nva + nvb
    rng_a = np.array(rng_, dtype=I64)
    offset = np.zeros(ntile, dtype=I64)
    offset[1:] = np.cumsum(rng_a)[:-1]
    alpha = np.arange(1, ntile + 1, dtype=I64)
    if restricted:  # tce_tile.F:1293-1306
        alpha[noa:noa + nob] = np.arange(1, noa + 1)
        alpha[noa + nob + nva:] = np.arange(noa + nob + 1, noa + nob + nva + 1)
    eps_spatial = np.concatenate([eo, ev])
    evl_sorted = np.concatenate([eps_spatial[m] for m in members]) if ntile else np.zeros(0)
    return Tiling(noab=noa + nob, nvab=nva + nvb, restricted=restricted, spin=np.array(spin, dtype=I64),
                  sym=np.array(sym, dtype=I64), range=rng_a, offset=offset, alpha=alpha,
                  evl_sorted=np.ascontiguousarray(evl_sorted, dtype=np.float64), members=members)


def _hash(keys: list[int], sizes: list[int]) -> tuple[np.ndarray, int]:
    """TCE offset table: [n, keys..., offsets...]; keys come out ascending by construction."""
    n = len(keys)
    h = np.zeros(2 * n + 1, dtype=I64)
    h[0] = n
    h[1:n + 1] = keys
    off = np.zeros(n, dtype=I64)
    if n:
        off[1:] = np.cumsum(np.array(sizes, dtype=I64))[:-1]
    h[n + 1:] = off
    assert np.all(np.diff(h[1:n + 1]) > 0), "TCE keys must be strictly ascending"
    return h, int(sum(sizes))


def t1_offset(t: Tiling, irrep: int = 0):
    keys, sizes = [], []
    for p5b in range(t.noab + 1, t.noab + t.nvab + 1):
        for h6b in range(1, t.noab + 1):
            if t.spin[p5b - 1] != t.spin[h6b - 1]:
                continue
            if (t.sym[p5b - 1] ^ t.sym[h6b - 1]) != irrep:
                continue
            if t.restricted and t.spin[p5b - 1] + t.spin[h6b - 1] == 4:
                continue
            keys.append(h6b - 1 + t.noab * (p5b - t.noab - 1))
            sizes.append(t.r(p5b) * t.r(h6b))
    return _hash(keys, sizes)


def t2_offset(t: Tiling, irrep: int = 0):
    keys, sizes = [], []
    sp, sy = t.spin, t.sym
    for p1b in range(t.noab + 1, t.noab + t.nvab + 1):
        for p2b in range(p1b, t.noab + t.nvab + 1):
            for h3b in range(1, t.noab + 1):
                for h4b in range(h3b, t.noab + 1):
                    if sp[p1b - 1] + sp[p2b - 1] != sp[h3b - 1] + sp[h4b - 1]:
                        continue
                    if (sy[p1b - 1] ^ sy[p2b - 1] ^ sy[h3b - 1] ^ sy[h4b - 1]) != irrep:
                        continue
                    if t.restricted and sp[p1b - 1] + sp[p2b - 1] + sp[h3b - 1] + sp[h4b - 1] == 8:
                        continue
                    keys.append(h4b - 1 + t.noab * (h3b - 1 + t.noab * (p2b - t.noab - 1 + t.nvab * (p1b - t.noab - 1))))
                    sizes.append(t.r(p1b) * t.r(p2b) * t.r(h3b) * t.r(h4b))
    return _hash(keys, sizes)


def v2_offset(t: Tiling, irrep_v: int = 0, triples_only: bool = True):
    """tce_mo2e_offset.F key order (g3b<=g4b, g1b<=g2b).  With triples_only the table holds only the
    three classes the (T) path reads: <pp||hh>, <hp||hh>, <pp||hp>."""
    keys, sizes = [], []
    sp, sy = t.spin, t.sym
    N = t.noab + t.nvab
    is_p = lambda b: b > t.noab
    for g3b in range(1, N + 1):
        for g4b in range(g3b, N + 1):
            for g1b in range(1, N + 1):
                for g2b in range(g1b, N + 1):
                    if triples_only:
                        cls = (is_p(g3b), is_p(g4b), is_p(g1b), is_p(g2b))
                        if cls not in ((True, True, False, False), (False, True, False, False), (True, True, False, True)):
                            continue
                    if sp[g3b - 1] + sp[g4b - 1] != sp[g1b - 1] + sp[g2b - 1]:
                        continue
                    if (sy[g3b - 1] ^ sy[g4b - 1] ^ sy[g1b - 1] ^ sy[g2b - 1]) != irrep_v:
                        continue
                    if t.restricted and sp[g3b - 1] + sp[g4b - 1] + sp[g1b - 1] + sp[g2b - 1] == 8:
                        continue
                    keys.append(g2b - 1 + N * (g1b - 1 + N * (g4b - 1 + N * (g3b - 1))))
                    sizes.append(t.r(g3b) * t.r(g4b) * t.r(g1b) * t.r(g2b))
    return _hash(keys, sizes)


# ---- Lambda-CCSD(T) inputs (src/tce/ccsd_t/lambda_ccsd_t_left.F): lambda_1, lambda_2 and the (h,p) Fock blocks ----
def y1_offset(t: Tiling, irrep: int = 0):
    """lambda_1 blocks (h4b, p1b), key p1b-noab-1 + nvab*(h4b-1) (lambda_ccsd_t_left.F:154-155)."""
    keys, sizes = [], []
    for h4b in range(1, t.noab + 1):
        for p1b in range(t.noab + 1, t.noab + t.nvab + 1):
            if t.spin[h4b - 1] != t.spin[p1b - 1] or (t.sym[h4b - 1] ^ t.sym[p1b - 1]) != irrep:
                continue
            if t.restricted and t.spin[h4b - 1] + t.spin[p1b - 1] == 4:
                continue
            keys.append(p1b - t.noab - 1 + t.nvab * (h4b - 1))
            sizes.append(t.r(h4b) * t.r(p1b))
    return _hash(keys, sizes)


def y2_offset(t: Tiling, irrep: int = 0):
    """lambda_2 blocks (h4b<=h5b, p1b<=p2b), key p2b-noab-1 + nvab*(p1b-noab-1 + nvab*(h5b-1 + noab*(h4b-1)))
    (lambda_ccsd_t_left.F:378-380)."""
    keys, sizes = [], []
    sp, sy = t.spin, t.sym
    for h4b in range(1, t.noab + 1):
        for h5b in range(h4b, t.noab + 1):
            for p1b in range(t.noab + 1, t.noab + t.nvab + 1):
                for p2b in range(p1b, t.noab + t.nvab + 1):
                    if sp[h4b - 1] + sp[h5b - 1] != sp[p1b - 1] + sp[p2b - 1]:
                        continue
                    if (sy[h4b - 1] ^ sy[h5b - 1] ^ sy[p1b - 1] ^ sy[p2b - 1]) != irrep:
                        continue
                    if t.restricted and sp[h4b - 1] + sp[h5b - 1] + sp[p1b - 1] + sp[p2b - 1] == 8:
                        continue
                    keys.append(p2b - t.noab - 1 + t.nvab * (p1b - t.noab - 1 + t.nvab * (h5b - 1 + t.noab * (h4b - 1))))
                    sizes.append(t.r(h4b) * t.r(h5b) * t.r(p1b) * t.r(p2b))
    return _hash(keys, sizes)


def f1_hp_offset(t: Tiling, irrep: int = 0):
    """The (hole, particle) blocks of the Fock file, key p3b-1 + (noab+nvab)*(h6b-1) (lambda_ccsd_t_left.F:390-391);
    the other classes of f1 are not read by the (T)-type corrections and are left out of the table."""
    keys, sizes = [], []
    N = t.noab + t.nvab
    for h6b in range(1, t.noab + 1):
        for p3b in range(t.noab + 1, N + 1):
            if t.spin[h6b - 1] != t.spin[p3b - 1] or (t.sym[h6b - 1] ^ t.sym[p3b - 1]) != irrep:
                continue
            if t.restricted and t.spin[h6b - 1] + t.spin[p3b - 1] == 4:
                continue
            keys.append(p3b - 1 + N * (h6b - 1))
            sizes.append(t.r(h6b) * t.r(p3b))
    return _hash(keys, sizes)


# ---- CR-CCSD(T) intermediates (src/tce/ccsd_t/cr_ccsd_t_N.F, cr_ccsd_t_E.F): the three block stores the per-tuple
#      routines cr_ccsd_t_N_1 / _N_2 / _E_2 read (the reference can also load them from files: read_in3, gr1_1/gr1_2/ei1_2)
def cr_n1_offset(t: Tiling, irrep_v: int = 0):
    """i1(h11 p4 h1 h2) of cr_ccsd_t_N_1, stored (p4b, h11b, h1b<=h2b), h2 fastest; key
    h2b-1 + noab*(h1b-1 + noab*(h11b-1 + noab*(p4b-noab-1)))  (OFFSET_cr_ccsd_t_N_1_1, cr_ccsd_t_N.F:773-841)."""
    keys, sizes = [], []
    sp, sy = t.spin, t.sym
    for p4b in range(t.noab + 1, t.noab + t.nvab + 1):
        for h11b in range(1, t.noab + 1):
            for h1b in range(1, t.noab + 1):
                for h2b in range(h1b, t.noab + 1):
                    if sp[h11b - 1] + sp[p4b - 1] != sp[h1b - 1] + sp[h2b - 1]:
                        continue
                    if (sy[h11b - 1] ^ sy[p4b - 1] ^ sy[h1b - 1] ^ sy[h2b - 1]) != irrep_v:
                        continue
                    if t.restricted and sp[h11b - 1] + sp[p4b - 1] + sp[h1b - 1] + sp[h2b - 1] == 8:
                        continue
                    keys.append(h2b - 1 + t.noab * (h1b - 1 + t.noab * (h11b - 1 + t.noab * (p4b - t.noab - 1))))
                    sizes.append(t.r(p4b) * t.r(h11b) * t.r(h1b) * t.r(h2b))
    return _hash(keys, sizes)


def cr_n2_offset(t: Tiling, irrep_v: int = 0):
    """i1(p4 p5 h1 p12) of cr_ccsd_t_N_2, stored (p4b<=p5b, h1b, p12b), p12 fastest; key
    p12b-noab-1 + nvab*(h1b-1 + noab*(p5b-noab-1 + nvab*(p4b-noab-1)))  (OFFSET_cr_ccsd_t_N_2_1, cr_ccsd_t_N.F:4011-4079)."""
    keys, sizes = [], []
    sp, sy = t.spin, t.sym
    for p4b in range(t.noab + 1, t.noab + t.nvab + 1):
        for p5b in range(p4b, t.noab + t.nvab + 1):
            for h1b in range(1, t.noab + 1):
                for p12b in range(t.noab + 1, t.noab + t.nvab + 1):
                    if sp[p4b - 1] + sp[p5b - 1] != sp[h1b - 1] + sp[p12b - 1]:
                        continue
                    if (sy[p4b - 1] ^ sy[p5b - 1] ^ sy[h1b - 1] ^ sy[p12b - 1]) != irrep_v:
                        continue
                    if t.restricted and sp[p4b - 1] + sp[p5b - 1] + sp[h1b - 1] + sp[p12b - 1] == 8:
                        continue
                    keys.append(p12b - t.noab - 1 + t.nvab * (h1b - 1 + t.noab * (p5b - t.noab - 1 + t.nvab * (p4b - t.noab - 1))))
                    sizes.append(t.r(p4b) * t.r(p5b) * t.r(h1b) * t.r(p12b))
    return _hash(keys, sizes)


def cr_e2_offset(t: Tiling):
    """i1(p4 p5 h1 h2)_tt of cr_ccsd_t_E_2: the T2 block structure with target irrep irrep_t xor irrep_t = 0
    (OFFSET_cr_ccsd_t_E_2_1, cr_ccsd_t_E.F:907-960)."""
    return t2_offset(t, 0)


def decode_cr_n1_key(t: Tiling, key: int):
    h2b = key % t.noab + 1; key //= t.noab
    h1b = key % t.noab + 1; key //= t.noab
    h11b = key % t.noab + 1; key //= t.noab
    return key + t.noab + 1, h11b, h1b, h2b  # (p4b,h11b,h1b,h2b)


def decode_cr_n2_key(t: Tiling, key: int):
    p12b = key % t.nvab + t.noab + 1; key //= t.nvab
    h1b = key % t.noab + 1; key //= t.noab
    p5b = key % t.nvab + t.noab + 1; key //= t.nvab
    return key + t.noab + 1, p5b, h1b, p12b  # (p4b,p5b,h1b,p12b)


def decode_t1_key(t: Tiling, key: int):
    return key // t.noab + t.noab + 1, key % t.noab + 1  # (p5b, h6b)


def decode_t2_key(t: Tiling, key: int):
    h4b = key % t.noab + 1; key //= t.noab
    h3b = key % t.noab + 1; key //= t.noab
    p2b = key % t.nvab + t.noab + 1; key //= t.nvab
    return key + t.noab + 1, p2b, h3b, h4b  # (p1b,p2b,h3b,h4b)


def decode_v2_key(t: Tiling, key: int):
    N = t.noab + t.nvab
    g2b = key % N + 1; key //= N
    g1b = key % N + 1; key //= N
    g4b = key % N + 1; key //= N
    return key + 1, g4b, g1b, g2b  # (g3b,g4b,g1b,g2b)


# ------------------------------------------------------------------------------------------------
# `2eorb` storage (intorb): V2 kept spin-free over the alpha tiles, SURVEY 8f-2
# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class AlphaTiling:
    """The alpha-space arrays tce_tile.F builds when intorb is set (tce_tile.F:1156-1212, :1376-1383):
    noa hole + nva particle tiles; b2am maps every spin-orbital tile (1-based) to its alpha-space tile."""
    noa: int
    nva: int
    b2am: np.ndarray          # [noab+nvab]
    spin_alpha: np.ndarray    # [noa+nva] (all 1 for a closed-shell reference)
    sym_alpha: np.ndarray
    range_alpha: np.ndarray
    members: list             # spatial orbital ids of each alpha tile


def alpha_tiling(t: Tiling) -> AlphaTiling:
    """No active tiles (activecalc off): hole beta j -> hole alpha j, particle beta j -> particle alpha j."""
    noa = int(np.sum(t.spin[:t.noab] == 1)); nob = t.noab - noa
    nva = int(np.sum(t.spin[t.noab:] == 1)); nvb = t.nvab - nva
    if (nob, nvb) != (noa, nva):
        raise ValueError("2eorb needs a closed-shell tiling (same alpha and beta tiles)")
    b2am = np.zeros(t.noab + t.nvab, dtype=I64)
    b2am[:noa] = np.arange(1, noa + 1)                         # hole alpha          (tce_tile.F:1162-1164)
    b2am[noa:t.noab] = np.arange(1, noa + 1)                   # hole beta           (:1166-1172)
    b2am[t.noab:t.noab + nva] = noa + np.arange(1, nva + 1)    # particle alpha      (:1174-1176)
    b2am[t.noab + nva:] = noa + np.arange(1, nva + 1)          # particle beta       (:1192-1207)
    idx = list(range(noa)) + list(range(t.noab, t.noab + nva))
    return AlphaTiling(noa, nva, b2am, np.ones(noa + nva, dtype=I64), t.sym[idx].copy(), t.range[idx].copy(),
                       [t.members[i] for i in idx])


def index_pair(i: int, j: int) -> int:   # tce_mo2e_offset_intorb.F:615
    return (i * (i - 1)) // 2 + j


def v2orb_blocks(a: AlphaTiling, irrep_v: int = 0):
    """Stored orbital blocks in storage order: (g3b, g4b, g1b, g2b, key, offset, size), tce_mo2e_offset_intorb.F:32-50."""
    N = a.noa + a.nva
    out, size = [], 0
    for g3b in range(1, N + 1):
        for g4b in range(g3b, N + 1):
            for g1b in range(1, N + 1):
                for g2b in range(g1b, N + 1):
                    if a.spin_alpha[g3b - 1] + a.spin_alpha[g4b - 1] != a.spin_alpha[g1b - 1] + a.spin_alpha[g2b - 1]:
                        continue
                    if (a.sym_alpha[g3b - 1] ^ a.sym_alpha[g4b - 1] ^ a.sym_alpha[g1b - 1] ^ a.sym_alpha[g2b - 1]) != irrep_v:
                        continue
                    if index_pair(g4b, g3b) < index_pair(g2b, g1b):
                        continue
                    n = int(a.range_alpha[g3b - 1] * a.range_alpha[g4b - 1] * a.range_alpha[g1b - 1] * a.range_alpha[g2b - 1])
                    key = g2b - 1 + N * (g1b - 1 + N * (g4b - 1 + N * (g3b - 1)))
                    out.append((g3b, g4b, g1b, g2b, key, size, n))
                    size += n
    return out, size


def v2orb_offset(a: AlphaTiling, irrep_v: int = 0, idiv2e: int = 2):
    """The checkpointed offset table `k_v2_alpha_offset` exactly as tce_mo2e_offset_intorb.F:52-150 lays it out:
    [length1 | keys(length1+1) | offsets | g3b | g4b | g1b | g2b], one entry per `ipiece_l` stored blocks
    (idiv2e = 2 in tce_energy.F:673); tce_hash_v2 walks the block loops from the nearest checkpoint."""
    blocks, size = v2orb_blocks(a, irrep_v)
    length = len(blocks)
    ipiece_l = length // idiv2e
    if ipiece_l * idiv2e == length:
        length1 = idiv2e
    else:
        length1 = idiv2e + 1
    if length == 0 or ipiece_l == 0:
        raise ValueError("too few orbital blocks for idiv2e checkpoints")
    tab = np.zeros(6 * (length1 + 1) + 1, dtype=I64)
    tab[0] = length1
    ipos, i_counter, addr = 1, 0, 0

    def put(pos, blk):
        g3b, g4b, g1b, g2b, key, off, _ = blk
        tab[pos] = key
        tab[(length1 + 1) + pos] = off
        tab[2 * (length1 + 1) + pos] = g3b
        tab[3 * (length1 + 1) + pos] = g4b
        tab[4 * (length1 + 1) + pos] = g1b
        tab[5 * (length1 + 1) + pos] = g2b

    last = None
    for blk in blocks:
        i_counter += 1
        if addr == 0:
            put(ipos, blk); ipos += 1
        if i_counter == ipiece_l:
            put(ipos, blk); ipos += 1
            i_counter = 0
        addr += 1
        last = blk
    if i_counter != 0:
        put(ipos, last)
    return tab, size
