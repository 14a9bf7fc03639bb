"""Timing of nwc_triples_run_creom (CR-EOMCCSD(T) tuple loop) on a named shape with random inputs:
python tools/creom_bench.py [shape] [max_tasks] [out.json].  Reports the (T) run of the same tasks beside it."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi, synth, tiling as tl
shape = sys.argv[1] if len(sys.argv) > 1 else "microbench_t40"
max_tasks = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = sys.argv[3] if len(sys.argv) > 3 else None
t = synth.shape_tiling(shape)
st = synth.random_blocks(t)
rng = np.random.default_rng(5)
n1h, n1 = tl.cr_n1_offset(t); n2h, n2 = tl.cr_n2_offset(t); e2h, e2 = tl.cr_e2_offset(t)
r = lambda n, s: rng.uniform(-1, 1, n) * s


class Obj:
    pass


cr = Obj()
cr.n1_hash, cr.n1, cr.n2_hash, cr.n2, cr.e2_hash, cr.e2 = n1h, r(n1, 0.1), n2h, r(n2, 0.1), e2h, r(e2, 0.02)
q = Obj()
q.x1_hash, q.x1, q.x2_hash, q.x2 = st.t1_hash, r(len(st.t1), 0.05), st.t2_hash, r(len(st.t2), 0.02)
q.m1_hash, q.m1, q.m2_hash, q.m2, q.m3_hash, q.m3, q.m4_hash, q.m4 = n1h, r(n1, 0.1), n2h, r(n2, 0.1), n1h, r(n1, 0.1), n2h, r(n2, 0.1)
q.q2_hash, q.q2, q.r0, q.excit = e2h, r(e2, 0.02), 0.4, 0.25
tr = capi.Triples(0)
tr.set_state(st)
tr.set_cr(cr)
tr.set_creom(q)
res = {}
def composed():
    os.environ["NWC_CREOM_COMPOSED"] = "1"
    try:
        return tr.run_creom(max_tasks=max_tasks)
    finally:
        del os.environ["NWC_CREOM_COMPOSED"]


for name, fn in (("(T)", lambda: tr.run(max_tasks=max_tasks)), ("CR-EOM-(T)", lambda: tr.run_creom(max_tasks=max_tasks)),
                 ("CR-EOM-(T) composed", composed)):
    fn()
    tr.set_timing(True); tr.stats(reset=True)
    t0 = time.time(); e = fn(); dt = time.time() - t0
    s = tr.stats()
    res[name] = dict(shape=shape, wall_s=dt, fused_ms=s["fused_ms"], flops=s["flops"], tflops=s["flops"] / dt * 1e-12,
                     launches=int(s["fused_launches"]), result=[float(x) for x in np.ravel(e)])
    print(f"{name:19s} {shape}: {dt:.3f} s wall, fused {s['fused_ms']:.1f} ms, executed {s['flops']:.3e} FLOP = "
          f"{s['flops'] / dt * 1e-12:.2f} TFLOP/s, launches {int(s['fused_launches'])}, result {e}", flush=True)
if out:
    json.dump(res, open(out, "w"), indent=1)
