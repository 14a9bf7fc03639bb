"""Timing of nwc_triples_run_lambda on a named shape with random lambda_1 / lambda_2 / Fock stores:
python tools/lambda_bench.py [shape] [max_tasks].  Reports the (T) run of the same tasks beside it."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi, synth, tiling as tl
shape = sys.argv[1] if len(sys.argv) > 1 else "microbench_t40"
max_tasks = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t = synth.shape_tiling(shape)
st = synth.random_blocks(t)
rng = np.random.default_rng(5)
y1h, n1 = tl.y1_offset(t); y2h, n2 = tl.y2_offset(t); f1h, nf = tl.f1_hp_offset(t)
lam = synth.LambdaStores(y1h, rng.uniform(-1, 1, n1) * 0.05, y2h, rng.uniform(-1, 1, n2) * 0.02, f1h, rng.uniform(-1, 1, nf) * 0.01)
tr = capi.Triples(0)
tr.set_state(st)
tr.set_lambda(lam)
for name, fn in (("(T)", lambda: tr.run(max_tasks=max_tasks)), ("Lambda-(T)", lambda: tr.run_lambda(max_tasks=max_tasks))):
    fn(); fn()
    tr.set_timing(True); tr.stats(reset=True)
    t0 = time.time(); e = fn(); dt = time.time() - t0
    s = tr.stats()
    print(f"{name:11s} {shape}: {dt:.3f} s wall, fused {s['fused_ms']:.1f} ms, executed {s['flops']:.3e} FLOP = "
          f"{s['flops'] / dt * 1e-12:.2f} TFLOP/s, repack {s['repack_ms']:.1f} ms, launches {int(s['fused_launches'])}, energies {e}", flush=True)
