import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi, synth
name = sys.argv[1] if len(sys.argv) > 1 else "microbench_t40"
t = synth.shape_tiling(name)
t0 = time.time(); st = synth.random_blocks(t); print("gen", time.time() - t0, "s", flush=True)
tr = capi.Triples(0)
if os.environ.get("INTORB"):   # `2eorb` storage: V2 antisymmetrised on the device from an orbital-form store
    t0 = time.time(); st.orb = synth.random_orbital(t); print("orbital store", len(st.orb.v2orb) * 8e-9, "GB", time.time() - t0, "s", flush=True)
    tr.set_state_2eorb(st)
else:
    tr.set_state(st)
tr.set_timing(True)
for it in range(int(os.environ.get("ITERS", "3"))):
    tr.stats(reset=True)
    t0 = time.time(); e1, e2 = tr.run(); dt = time.time() - t0
    s = tr.stats()
    print(json.dumps(dict(it=it, e1=e1, e2=e2, wall_s=dt, fused_ms=s["fused_ms"], repack_ms=s["repack_ms"], flops=s["flops"],
                          tflops_wall=s["flops"] / dt * 1e-12, tflops_fused=s["flops"] / s["fused_ms"] * 1e-9,
                          items=s["work_items"], descs=s["descs"])), flush=True)
