import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
st = synth.random_blocks(synth.shape_tiling("microbench_t40"))
tr = capi.Triples(0); tr.set_state(st)
ref = None; bad = 0
for i in range(n):
    e1, e2, pt = tr.run(per_task=True)
    if ref is None: ref = (e1, e2, pt.copy())
    elif (e1, e2) != ref[:2]:
        bad += 1; print("MISMATCH run", i, e1 - ref[0], e2 - ref[1], (pt - ref[2]).tolist(), flush=True)
print("runs", n, "mismatches", bad, "ref", ref[0], ref[1])
