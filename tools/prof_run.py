"""One fused launch of a named shape for ncu: python tools/prof_run.py <shape> <first_task> <ntasks> [intorb]
Stores generated on the device; runs the given tasks of the heaviest-first list once (after one warm-up run)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi, synth
shape = sys.argv[1]; first = int(sys.argv[2]); n = int(sys.argv[3]); intorb = len(sys.argv) > 4 and sys.argv[4] == "intorb"
t = synth.shape_tiling(shape)
tr = capi.Triples(0)
st = synth.empty_stores(t, intorb=intorb)
if intorb:
    tr.set_state_2eorb(st)
else:
    tr.set_state(st)
tr.synth_fill(20240229, (1e-3, 5e-5, 5e-3))
tr.set_batch_bytes(64 << 30)
ids = np.arange(first, first + n, dtype=np.int64) if n > 0 else np.array([first], np.int64)
for _ in range(2):
    tr.stats(reset=True)
    tr.set_timing(True)
    e = tr.run_partition_list(0, 1, ids)
    s = tr.stats()
    print(shape, ids.tolist(), e, "fused_ms", s["fused_ms"], "TF", s["flops"] / max(s["fused_ms"], 1e-9) * 1e-9, "order", tr.order, flush=True)
