"""Per-CTA phase clocks of the fused kernel (debug build path): python tools/phase_timing.py [shape]"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi, synth
name = sys.argv[1] if len(sys.argv) > 1 else "microbench_t40"
st = synth.random_blocks(synth.shape_tiling(name))
tr = capi.Triples(0); tr.set_state(st)
tr.run(max_tasks=1)
cap = 400000
l = capi.lib()
l.nwc_debug_phase_timing(None, C.c_uint(cap), 0)
tr.set_timing(True); tr.stats(reset=True)
tr.run(max_tasks=1)
s = tr.stats()
buf = np.zeros((cap, 8), np.uint64)
l.nwc_debug_phase_timing(buf.ctypes.data_as(C.c_void_p), C.c_uint(cap), 1)
l.nwc_debug_phase_timing(None, C.c_uint(0), 0)
b = buf[2000:cap - 2000].astype(np.float64)   # skip the first/last waves
b = b[b[:, 6] > 0]
names = ["setup", "Kloops+xfers", "cta-barrier", "singles", "energy+reduce"]
waitc = b[:, 2].copy()
b = np.delete(b, 2, axis=1)
d = np.diff(b[:, :6], axis=1)
tot = b[:, 5] - b[:, 0]
print(f"{name}: CTAs sampled {len(b)}, fused_ms {s['fused_ms']:.1f}, mean CTA cycles {tot.mean():.0f}")
for i, n in enumerate(names):
    print(f"  {n:16s} {d[:, i].mean():9.0f} cyc  {d[:, i].mean() / tot.mean() * 100:5.1f}%   (p50 {np.median(d[:, i]):.0f})")
print(f"  canon<->acc transfers (inside K loops) {b[:, 6].mean():9.0f} cyc  {b[:, 6].mean() / tot.mean() * 100:5.1f}%")
print(f"  waiting for operands (inside K loops)  {waitc.mean():9.0f} cyc  {waitc.mean() / tot.mean() * 100:5.1f}%")
if os.environ.get("NB"):   # NB=6,6,6,10,10,10 : blocks per index (h3,h2,h1,p6,p5,p4) of the first task -> group by edge count
    nb = [int(x) for x in os.environ["NB"].split(",")]
    idx = np.arange(cap)[2000:cap - 2000]
    idx = idx[buf[2000:cap - 2000, 6] > 0]
    nedge = np.zeros(len(idx), int)
    r = idx.copy()
    for q in range(6):
        nedge += (r % nb[q] == nb[q] - 1)
        r //= nb[q]
    for e in range(7):
        m = nedge == e
        if m.any():
            print(f"  sub-tiles with {e} edge indices: {m.sum():7d}   mean K-loop cycles {d[m, 1].mean():9.0f}   wait {waitc[m].mean():9.0f}")
