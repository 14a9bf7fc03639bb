"""Bitwise run-to-run determinism of the fused kernel: python tools/stress.py [runs] [shape] [max_tasks].
A data race in the operand ring shows up here as energy noise."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
shape = sys.argv[2] if len(sys.argv) > 2 else "microbench_t40"
max_tasks = int(sys.argv[3]) if len(sys.argv) > 3 else 0
st = synth.random_blocks(synth.shape_tiling(shape))
tr = capi.Triples(0); tr.set_state(st)
ref = None; bad = 0
for i in range(n):
    e1, e2, pt = tr.run(per_task=True, max_tasks=max_tasks)
    if ref is None: ref = (e1, e2, pt.copy())
    elif (e1, e2) != ref[:2] or not np.array_equal(pt, ref[2]):
        bad += 1; print("MISMATCH run", i, e1 - ref[0], e2 - ref[1], np.abs(pt - ref[2]).max(), flush=True)
print("shape", shape, "runs", n, "mismatches", bad, "ref", ref[0], ref[1])
