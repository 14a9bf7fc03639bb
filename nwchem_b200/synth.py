"""Synthetic T1/T2/V2 block stores of a given molecular shape (there is no SCF/CCSD stack in scope).

Two generators (SURVEY 7.2 / 8d "concrete synthetic inputs"):
  * random_blocks  -- every stored block iid U(-1,1)*scale, seeded per block key (microbench, large shapes)
  * physical       -- spatial t1/t2/(pq|rs) with the right permutational symmetry, spin-integrated and
                      antisymmetrised into TCE spin-orbital blocks; gives tile-size-invariant E[T]/E(T).
Block element order is TCE's: indices in key order, LAST index fastest (src/tce/sort/tce_sort4.F:31).
"""
from __future__ import annotations
import dataclasses
import numpy as np
from . import tiling as tl


@dataclasses.dataclass
class OrbitalV2:
    """`2eorb` storage of the two-electron integrals (tce intorb): spin-free blocks over the alpha tiles."""
    a: tl.AlphaTiling
    v2orb_hash: np.ndarray    # k_v2_alpha_offset (checkpointed table, tce_mo2e_offset_intorb.F)
    v2orb: np.ndarray         # d_v2orb


@dataclasses.dataclass
class BlockStores:
    t: tl.Tiling
    t1_hash: np.ndarray; t1: np.ndarray
    t2_hash: np.ndarray; t2: np.ndarray
    v2_hash: np.ndarray; v2: np.ndarray
    orb: OrbitalV2 | None = None   # when set, V2 is read from the orbital-form store instead of v2_hash/v2


def _iter_hash(h):
    n = int(h[0])
    for i in range(n):
        yield int(h[1 + i]), int(h[1 + n + i])


def random_blocks(t: tl.Tiling, seed: int = 20240229, scale=(0.05, 0.02, 0.1)) -> BlockStores:
    t1h, n1 = tl.t1_offset(t); t2h, n2 = tl.t2_offset(t); v2h, nv = tl.v2_offset(t)
    rng = np.random.default_rng(seed)
    t1 = rng.uniform(-1, 1, n1) * scale[0]
    t2 = rng.uniform(-1, 1, n2) * scale[1]
    v2 = rng.uniform(-1, 1, nv) * scale[2]
    return BlockStores(t, t1h, t1, t2h, t2, v2h, v2)


# ---- the keyed generator of nwc_triples_synth_fill (csrc/kernels.cu synth_fill_kernel), restated in numpy ----
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_C1, _C2, _C3 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xD1B54A32D192ED03), np.uint64(0xBF58476D1CE4E5B9)
_C4 = np.uint64(0x94D049BB133111EB)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * _C3
    z = (z ^ (z >> np.uint64(27))) * _C4
    return z ^ (z >> np.uint64(31))


def keyed_values(seed: int, store: int, key: int, n: int, scale: float) -> np.ndarray:
    """Element e of block `key` of store `store` (1 T1, 2 T2, 3 spin-orbital V2, 4 orbital V2): scale*(2u-1),
    u = top 53 bits of a splitmix64-style hash of (seed, store, key, e).  Bit-identical to the device generator."""
    with np.errstate(over="ignore"):
        hk = _mix(np.array([(np.uint64(seed) * _C1 + np.uint64(store) * _C2) ^ np.uint64(key)], np.uint64))
        h = _mix(hk + np.arange(n, dtype=np.uint64) * _C1)
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return scale * (2.0 * u - 1.0)


def _fill_table(h, total, seed, store, scale):
    out = np.zeros(total)
    n = int(h[0])
    offs = [int(h[1 + n + i]) for i in range(n)] + [total]
    for i in range(n):
        out[offs[i]:offs[i + 1]] = keyed_values(seed, store, int(h[1 + i]), offs[i + 1] - offs[i], scale)
    return out


def keyed_blocks(t: tl.Tiling, seed: int = 20240229, scale=(0.05, 0.02, 0.1), intorb: bool = False) -> BlockStores:
    """Host copy of what Triples.synth_fill(seed, scale) puts on the device (small shapes: oracle parity)."""
    t1h, n1 = tl.t1_offset(t); t2h, n2 = tl.t2_offset(t); v2h, nv = tl.v2_offset(t)
    st = BlockStores(t, t1h, _fill_table(t1h, n1, seed, 1, scale[0]), t2h, _fill_table(t2h, n2, seed, 2, scale[1]),
                     v2h, None if intorb else _fill_table(v2h, nv, seed, 3, scale[2]))
    if intorb:
        a = tl.alpha_tiling(t)
        tab, size = tl.v2orb_offset(a)
        blocks, _ = tl.v2orb_blocks(a)
        vo = np.zeros(size)
        is_p = lambda b: b > a.noa
        for g3b, g4b, g1b, g2b, key, off, n in blocks:
            # the device holds (and fills) only the blocks (T) can touch: one mixed hole/particle tile pair at least
            if int(is_p(g3b)) + int(is_p(g4b)) == 1 or int(is_p(g1b)) + int(is_p(g2b)) == 1:
                vo[off:off + n] = keyed_values(seed, 4, key, n, scale[2])
        st.orb = OrbitalV2(a, tab, vo)
        st.v2 = np.zeros(0)
    return st


def empty_stores(t: tl.Tiling, intorb: bool = False) -> BlockStores:
    """Tables only, no data: the library allocates the stores and Triples.synth_fill generates them on the device."""
    t1h, _ = tl.t1_offset(t); t2h, _ = tl.t2_offset(t)
    if intorb:
        a = tl.alpha_tiling(t)
        tab, _ = tl.v2orb_offset(a)
        return BlockStores(t, t1h, None, t2h, None, np.zeros(1, np.int64), None, OrbitalV2(a, tab, None))
    v2h, _ = tl.v2_offset(t)
    return BlockStores(t, t1h, None, t2h, None, v2h, None)


def random_orbital(t: tl.Tiling, seed: int = 20240229, scale: float = 0.1) -> OrbitalV2:
    """iid orbital-form V2 store for large shapes (timing, oracle-vs-GPU parity); not tied to a spin-orbital store."""
    a = tl.alpha_tiling(t)
    tab, size = tl.v2orb_offset(a)
    rng = np.random.default_rng(seed + 1)
    return OrbitalV2(a, tab, rng.uniform(-1, 1, size) * scale)


def physical_dense(t: tl.Tiling, seed: int = 20240229, naux: int = 12):
    """The dense closed-shell spatial tensors behind `physical`: (no, nv, t1s[a,i], t2s[a,b,i,j], eri[p,q,r,s] = (pq|rs)).
    Drawn in a tiling-independent order, so different tilesizes (and the dense spin-orbital checks of the oracle) see
    identical tensors."""
    no = int(sum(t.range[i] for i in range(t.noab) if t.spin[i] == 1))
    nv = int(sum(t.range[i] for i in range(t.noab, t.noab + t.nvab) if t.spin[i] == 1))
    n = no + nv
    # irrep of each spatial orbital (from the alpha tiles)
    irr = np.zeros(n, dtype=np.int64)
    for b in range(t.noab + t.nvab):
        if t.spin[b] == 1:
            irr[t.members[b]] = t.sym[b]
    rng = np.random.default_rng(seed)
    # NOTE: draw in a tiling-independent order so different tilesizes see identical tensors
    t1s = rng.uniform(-1, 1, (nv, no)) * 0.05
    t2s = rng.uniform(-1, 1, (nv, nv, no, no)) * 0.02
    t2s = 0.5 * (t2s + t2s.transpose(1, 0, 3, 2))
    B = rng.uniform(-1, 1, (naux, n, n)) * 0.3
    B = 0.5 * (B + B.transpose(0, 2, 1))
    gam = rng.integers(0, int(irr.max()) + 1, naux)
    pair = irr[:, None] ^ irr[None, :]
    for L in range(naux):
        B[L][pair != gam[L]] = 0.0
    eri = np.einsum("Lpq,Lrs->pqrs", B, B)  # (pq|rs), 8-fold symmetric, irrep-0
    io, iv = irr[:no], irr[no:]
    t1s[(iv[:, None] ^ io[None, :]) != 0] = 0.0
    m = iv[:, None, None, None] ^ iv[None, :, None, None] ^ io[None, None, :, None] ^ io[None, None, None, :]
    t2s[m != 0] = 0.0
    return no, nv, t1s, t2s, eri


def physical(t: tl.Tiling, seed: int = 20240229, naux: int = 12, intorb: bool = False, dense=None) -> BlockStores:
    """Dense spatial tensors -> spin-orbital antisymmetrised blocks.  Only for small orbital counts.
    With an unrestricted tiling (t.restricted False) the same closed-shell tensors are expanded into every spin
    block (beta tiles are their own owners), so E[T]/E(T) must equal the restricted result: a check of the
    `restricted` factor-2 / k_alpha logic (ccsd_t_dot.F:52-56, tce_restricted.F).
    dense = (no, nv, t1s[a,i], t2s[a,b,i,j], eri[p,q,r,s]) replaces the synthetic tensors (real CCSD amplitudes and MO
    integrals: oracle/h2o_ccsd.py), in the spatial-orbital order of the tiling (occupied then virtual)."""
    no, nv, t1s, t2s, eri = physical_dense(t, seed, naux) if dense is None else dense

    def so(b):  # spatial ids and spin of tile b (1-based)
        return t.members[b - 1], int(t.spin[b - 1])

    def v_block(g3b, g4b, g1b, g2b):
        (a, sa), (b, sb), (c, sc), (d, sd) = so(g3b), so(g4b), so(g1b), so(g2b)
        out = np.zeros((len(a), len(b), len(c), len(d)))
        if sa == sc and sb == sd:  # <ab|cd> = (ac|bd)
            out += eri[np.ix_(a, c, b, d)].transpose(0, 2, 1, 3)
        if sa == sd and sb == sc:  # <ab|dc> = (ad|bc)
            out -= eri[np.ix_(a, d, b, c)].transpose(0, 2, 3, 1)
        return out

    def t2_block(p1b, p2b, h3b, h4b):
        (a, sa), (b, sb), (i, si), (j, sj) = so(p1b), so(p2b), so(h3b), so(h4b)
        a, b = a - no, b - no
        out = np.zeros((len(a), len(b), len(i), len(j)))
        if sa == si and sb == sj:
            out += t2s[np.ix_(a, b, i, j)]
        if sa == sj and sb == si:
            out -= t2s[np.ix_(a, b, j, i)].transpose(0, 1, 3, 2)
        return out

    t1h, n1 = tl.t1_offset(t); t2h, n2 = tl.t2_offset(t); v2h, nv2 = tl.v2_offset(t)
    t1 = np.zeros(n1); t2 = np.zeros(n2); v2 = np.zeros(nv2)
    for key, off in _iter_hash(t1h):
        p5b, h6b = tl.decode_t1_key(t, key)
        (a, _), (i, _) = so(p5b), so(h6b)
        blk = t1s[np.ix_(a - no, i)]
        t1[off:off + blk.size] = blk.ravel()
    for key, off in _iter_hash(t2h):
        blk = t2_block(*tl.decode_t2_key(t, key))
        t2[off:off + blk.size] = blk.ravel()
    for key, off in _iter_hash(v2h):
        blk = v_block(*tl.decode_v2_key(t, key))
        v2[off:off + blk.size] = blk.ravel()
    orb = None
    if intorb:
        # the same integrals in `2eorb` form: block (g3b<=g4b | g1b<=g2b) over alpha tiles holds (k l|i j) with
        # k in g4b fastest, then l in g3b, i in g2b, j in g1b (tce_mo2e_trans.F:707-723)
        a = tl.alpha_tiling(t)
        tab, size = tl.v2orb_offset(a)
        blocks, _ = tl.v2orb_blocks(a)
        vo = np.zeros(size)
        for g3b, g4b, g1b, g2b, key, off, n in blocks:
            mk, ml, mi, mj = a.members[g4b - 1], a.members[g3b - 1], a.members[g2b - 1], a.members[g1b - 1]
            blk = eri[np.ix_(mk, ml, mi, mj)].transpose(3, 2, 1, 0)   # [j][i][l][k], k fastest
            vo[off:off + n] = blk.ravel()
        orb = OrbitalV2(a, tab, vo)
    return BlockStores(t, t1h, t1, t2h, t2, v2h, v2, orb)


@dataclasses.dataclass
class LambdaStores:
    """Inputs of the Lambda-CCSD(T) left-hand side: lambda_1 (h,p), lambda_2 (hh,pp), Fock (h,p) blocks."""
    y1_hash: np.ndarray; y1: np.ndarray
    y2_hash: np.ndarray; y2: np.ndarray
    f1_hash: np.ndarray; f1: np.ndarray


def physical_lambda(t: tl.Tiling, seed: int = 777) -> LambdaStores:
    """Closed-shell spatial lambda_1[i,a], lambda_2[i,j,a,b] (= lambda_2[j,i,b,a]) and f[i,a], spin-integrated into TCE
    blocks exactly like `physical` does for t1/t2 (tiling-independent draws, so tile-size invariance can be tested)."""
    no = int(sum(t.range[i] for i in range(t.noab) if t.spin[i] == 1))
    nv = int(sum(t.range[i] for i in range(t.noab, t.noab + t.nvab) if t.spin[i] == 1))
    irr = np.zeros(no + nv, dtype=np.int64)
    for b in range(t.noab + t.nvab):
        if t.spin[b] == 1:
            irr[t.members[b]] = t.sym[b]
    rng = np.random.default_rng(seed)
    y1s = rng.uniform(-1, 1, (no, nv)) * 0.05
    y2s = rng.uniform(-1, 1, (no, no, nv, nv)) * 0.02
    y2s = 0.5 * (y2s + y2s.transpose(1, 0, 3, 2))
    fs = rng.uniform(-1, 1, (no, nv)) * 0.01
    io, iv = irr[:no], irr[no:]
    y1s[(io[:, None] ^ iv[None, :]) != 0] = 0.0
    fs[(io[:, None] ^ iv[None, :]) != 0] = 0.0
    m = io[:, None, None, None] ^ io[None, :, None, None] ^ iv[None, None, :, None] ^ iv[None, None, None, :]
    y2s[m != 0] = 0.0

    def so(b):
        return t.members[b - 1], int(t.spin[b - 1])

    y1h, n1 = tl.y1_offset(t); y2h, n2 = tl.y2_offset(t); f1h, nf = tl.f1_hp_offset(t)
    y1 = np.zeros(n1); y2 = np.zeros(n2); f1 = np.zeros(nf)
    N = t.noab + t.nvab
    for key, off in _iter_hash(y1h):
        h4b, p1b = key // t.nvab + 1, key % t.nvab + t.noab + 1
        (i, _), (a, _) = so(h4b), so(p1b)
        blk = y1s[np.ix_(i, a - no)]
        y1[off:off + blk.size] = blk.ravel()
    for key, off in _iter_hash(f1h):
        h6b, p3b = key // N + 1, key % N + 1
        (i, _), (a, _) = so(h6b), so(p3b)
        blk = fs[np.ix_(i, a - no)]
        f1[off:off + blk.size] = blk.ravel()
    for key, off in _iter_hash(y2h):
        k = key
        p2b = k % t.nvab + t.noab + 1; k //= t.nvab
        p1b = k % t.nvab + t.noab + 1; k //= t.nvab
        h5b = k % t.noab + 1; k //= t.noab
        h4b = k + 1
        (i, si), (j, sj), (a, sa), (b, sb) = so(h4b), so(h5b), so(p1b), so(p2b)
        a, b = a - no, b - no
        blk = np.zeros((len(i), len(j), len(a), len(b)))
        if si == sa and sj == sb:
            blk += y2s[np.ix_(i, j, a, b)]
        if si == sb and sj == sa:
            blk -= y2s[np.ix_(i, j, b, a)].transpose(0, 1, 3, 2)
        y2[off:off + blk.size] = blk.ravel()
    return LambdaStores(y1h, y1, y2h, y2, f1h, f1)


# ---- named shapes of BASELINE.json's configs (alpha occ / alpha virt, C1 unless stated) ----
SHAPES = {
    # H2O cc-pVDZ on the exact QA tile table (tce_ccsd_t_h2o.out:644-659): irreps a1,a2,b1,b2 = 0,1,2,3
    "h2o_ccpvdz_c2v": dict(occ=[3, 0, 1, 1], virt=[8, 2, 6, 3], tilesize=20),
    "h2o_ccpvdz_c1": dict(occ=[5], virt=[19], tilesize=20),
    "microbench_t40": dict(occ=[40], virt=[40], tilesize=40),
    "uracil_augccpvdz": dict(occ=[21], virt=[191], tilesize=40),
    "benzene_dimer_augccpvtz": dict(occ=[30], virt=[786], tilesize=40),
    "h2o10_augccpvtz": dict(occ=[40], virt=[870], tilesize=40),
    # profiling stand-in for the (H2O)10 shape: same 40-wide tiles, 4 instead of 22 virtual tiles per spin, so the
    # resident store is < 1 GB (ncu's kernel replay saves and restores device memory around every pass)
    "h2o10_slice_v160": dict(occ=[40], virt=[160], tilesize=40),
}


def shape_tiling(name: str, tilesize: int | None = None, seed: int = 20240229, restricted: bool = True) -> tl.Tiling:
    s = SHAPES[name]
    return tl.make_tiling(s["occ"], s["virt"], tilesize or s["tilesize"], restricted, seed)


def shard_v2(st: BlockStores, rank: int, world: int) -> BlockStores:
    """The V2 shard of `rank`: block i of the offset table belongs to rank i % world; the shard keeps its blocks in
    table order, compacted.  Offset tables stay the full ones (every rank derives the same shard offsets)."""
    h = st.v2_hash
    n = int(h[0])
    offs = [int(h[1 + n + i]) for i in range(n)] + [len(st.v2)]
    parts = [st.v2[offs[i]:offs[i + 1]] for i in range(n) if i % world == rank]
    v2 = np.concatenate(parts) if parts else np.zeros(0)
    return BlockStores(st.t, st.t1_hash, st.t1, st.t2_hash, st.t2, st.v2_hash, np.ascontiguousarray(v2))


def shard_store(hash_table, data, rank: int, world: int):
    """The shard of `rank` of any block store: block i of the offset table belongs to rank i % world; the shard keeps its
    blocks in table order, compacted (nwc_triples_set_cr_sharded for the CR-CCSD(T) pphp intermediate)."""
    n = int(hash_table[0])
    offs = [int(hash_table[1 + n + i]) for i in range(n)] + [len(data)]
    parts = [data[offs[i]:offs[i + 1]] for i in range(n) if i % world == rank]
    return np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0)
