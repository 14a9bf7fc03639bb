"""CPU tests of the product's host side: the C-ABI library loads, exports every declared symbol, and its
host-side driver logic (task list, dispatch walk) agrees with the oracle.  No compute calls (no GPU here)."""
import os
import re
import dataclasses
import numpy as np
import pytest
from nwchem_b200 import capi, synth, tiling as tl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nwc_triples.h")).read()
    names = set(re.findall(r"\b(nwc_[a-z0-9_]+|check_device_|device_init_|initmemmodule_|finalizememmodule_|"
                           r"dev_mem_s_|dev_mem_d_|dev_release_|compute_en_)\s*\(", src))
    names -= {"nwc_tce_state", "nwc_triples_ctx", "nwc_triples_stats"}
    for fam in ("s1", "d1", "d2"):
        for k in range(1, 10):
            names.add(f"sd_t_{fam}_{k}_cuda_")
    return sorted(names)


def test_library_loads_and_exports_all_declared_symbols():
    l = capi.lib()
    syms = _declared_symbols()
    assert len(syms) >= 27 + 8 + 15
    for s in syms:
        assert hasattr(l, s), f"libnwc_triples.so does not export {s}"


def test_kernel_tables_match_between_python_and_c_header():
    from nwchem_b200.kernel_tables import DECL, SIGN, DECL_E1, SIGN_E1
    src = open(os.path.join(ROOT, "nwchem_b200", "csrc", "tables.h")).read()
    rows = re.findall(r"\{(N_[HP]\d, N_[HP]\d, N_[HP]\d, N_[HP]\d, N_[HP]\d, N_[HP]\d)\}", src)
    assert len(rows) == 36          # 27 (T) entry points + the nine sd_E_K of CR-CCSD(T)
    flat = [tuple(x.strip()[2:].lower() for x in r.split(",")) for r in rows]
    assert flat == [d for fam in DECL for d in fam] + list(DECL_E1)
    signs = re.findall(r"\{([+-]1(?:, [+-]1){8})\}", src)
    assert [tuple(int(x) for x in s.split(",")) for s in signs] == [tuple(s) for s in SIGN] + [tuple(SIGN_E1)]


@pytest.mark.parametrize("shape,ts", [("h2o_ccpvdz_c2v", 20), ("h2o_ccpvdz_c2v", 5), ("h2o_ccpvdz_c1", 7),
                                      ("uracil_augccpvdz", 40)])
def test_host_task_list_and_dispatch_match_oracle(oracle, shape, ts):
    t = synth.shape_tiling(shape, tilesize=ts)

    class D:
        pass
    d = D(); d.t = t
    d.t1_hash = d.t2_hash = d.v2_hash = np.zeros(3, np.int64); d.t1 = d.t2 = d.v2 = np.zeros(1)
    kl = capi.host_task_list(d)
    ko = oracle.task_list(t)
    assert np.array_equal(kl, ko)
    c, keep = oracle.make_ctx(d)
    state = capi.make_state(d)
    for tup in kl[:: max(1, len(kl) // 60)]:
        calls, flops = capi.host_count_tuple(d, tup[:6], state)
        cnt = oracle.count_tuple(c, tup[:6], keep)
        assert tuple(calls) == (cnt.calls_s1, cnt.calls_d1, cnt.calls_d2)
        assert tuple(flops) == (cnt.flops_s1, cnt.flops_d1, cnt.flops_d2)


def test_h2o10_flops_from_product_host_logic():
    t = synth.shape_tiling("h2o10_augccpvtz")

    class D:
        pass
    d = D(); d.t = t
    d.t1_hash = d.t2_hash = d.v2_hash = np.zeros(3, np.int64); d.t1 = d.t2 = d.v2 = np.zeros(1)
    kl = capi.host_task_list(d)
    assert len(kl) == 7590
    state = capi.make_state(d)
    tot = np.zeros(3); calls = np.zeros(3, np.int64)
    for tup in kl:
        c, f = capi.host_count_tuple(d, tup[:6], state)
        tot += f; calls += c
    assert tuple(calls) == (46046, 62744, 1380368)
    assert abs(tot.sum() - 4.52e17) / 4.52e17 < 5e-3


@pytest.mark.parametrize("restricted", [True, False])
def test_2eorb_host_plan_reproduces_spin_orbital_blocks(restricted):
    """The library's host logic for `2eorb` storage (which stored orbital block, which strides, which sign -- the
    device kernel only executes this plan): applied with numpy to the orbital-form store it must rebuild every
    spin-orbital V2 block the (T) path reads, bit for bit (get_block_ind.F:818-1538 collapsed into strides)."""
    t = synth.shape_tiling("h2o_ccpvdz_c2v", restricted=restricted)
    st = synth.physical(t, intorb=True)
    vo = st.orb.v2orb
    halves = set()
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
        halves.add((oa >= 0, ob >= 0))
        n = int(np.prod(dims))
        assert np.array_equal(blk, st.v2[off:off + n]), (g3b, g4b, g1b, g2b)
    assert (True, True) in halves and (True, False) in halves
    if not restricted:
        assert (False, True) in halves
    bad = dataclasses.replace(st, orb=dataclasses.replace(st.orb, v2orb_hash=st.orb.v2orb_hash.copy()))
    bad.orb.v2orb_hash[1] += 1
    with pytest.raises(RuntimeError):
        capi.host_2eorb_plan(bad, 4, 4, 1, 1)


def test_public_struct_layouts_match_the_ctypes_mirror(tmp_path):
    """include/nwc_triples.h compiled as plain C (it is the C ABI a Fortran/C caller sees): the sizes and field offsets
    of the public structs must be what nwchem_b200/capi.py mirrors with ctypes."""
    import ctypes as C
    import subprocess
    src = tmp_path / "layout.c"
    fields = {"nwc_tce_state": [f for f, _ in capi.TceState._fields_],
              "nwc_triples_stats": [f for f, _ in capi.Stats._fields_],
              "nwc_tce_orb_state": [f for f, _ in capi.OrbState._fields_]}
    body = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "nwc_triples.h")}"',
            'int main(void) {']
    for st_, fl in fields.items():
        body.append(f'  printf("{st_} %zu", sizeof({st_}));')
        for f in fl:
            body.append(f'  printf(" %zu", offsetof({st_}, {f}));')
        body.append('  printf("\\n");')
    body += ['  return 0;', '}']
    src.write_text("\n".join(body))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    mirror = {"nwc_tce_state": capi.TceState, "nwc_triples_stats": capi.Stats, "nwc_tce_orb_state": capi.OrbState}
    for line in out:
        if not line.strip():
            continue
        name, size, *offs = line.split()
        cls = mirror[name]
        assert int(size) == C.sizeof(cls), name
        assert [int(o) for o in offs] == [getattr(cls, f).offset for f, _ in cls._fields_], name


def test_fortran_binding_names_every_symbol_the_library_exports():
    """integration/nwc_triples_mod.F90 (the ISO_C_BINDING module a maintainer adds) cannot be compiled here (no Fortran
    compiler), but every bind(C, name=...) in it must be a symbol libnwc_triples.so exports, with all 27 kernel entry
    points present; and every function the public header declares must be exported too."""
    import re
    import subprocess
    mod = open(os.path.join(ROOT, "integration", "nwc_triples_mod.F90")).read()
    names = set(re.findall(r"bind\(C,\s*name='([A-Za-z0-9_]+)'\)", mod))
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert names <= exported, sorted(names - exported)
    for fam in ("s1", "d1", "d2"):
        for k in range(1, 10):
            assert f"sd_t_{fam}_{k}_cuda_" in names
    hdr = open(os.path.join(ROOT, "include", "nwc_triples.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nwc_[a-z0-9_]+|check_device_|device_init_|initmemmodule_|finalizememmodule_|dev_mem_s_|dev_mem_d_|dev_release_|compute_en_)\s*\(", hdr))
    declared -= {"nwc_tce_state", "nwc_tce_orb_state", "nwc_triples_stats", "nwc_triples_ctx"}
    assert declared <= exported, sorted(declared - exported)


def test_host_sort4_matches_oracle_for_all_permutations(oracle):
    """The host driver's cache-blocked TCE_SORT_4 against the oracle's plain restatement: all 24 permutations, ragged
    dims that are not multiples of the 16-wide tile."""
    import ctypes as C
    import itertools
    l = capi.lib()
    PD = C.POINTER(C.c_double)
    l.nwc_host_sort4.argtypes = [PD, PD, C.c_long, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
    ol = oracle.lib()
    ol.ora_tce_sort_4.argtypes = [PD, PD, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, C.c_double]
    rng = np.random.default_rng(4)
    for dims in ((5, 17, 3, 21), (40, 39, 2, 18), (1, 1, 33, 1)):
        x = rng.standard_normal(int(np.prod(dims)))
        for perm in itertools.permutations((1, 2, 3, 4)):
            a = np.zeros_like(x); b = np.zeros_like(x)
            l.nwc_host_sort4(x.ctypes.data_as(PD), a.ctypes.data_as(PD), *dims, *perm, -0.5)
            ol.ora_tce_sort_4(x.ctypes.data_as(PD), b.ctypes.data_as(PD), *dims, *perm, -0.5)
            assert np.array_equal(a, b), (dims, perm)
