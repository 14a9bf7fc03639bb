"""One CR-CCSD(T) launch of a named shape for ncu: python tools/cr_prof_run.py <shape> <ntasks>
(random stores and intermediates; one warm-up run, then the measured one -- profile with --launch-skip 1)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi, synth, tiling as tl
shape = sys.argv[1]; n = int(sys.argv[2])
t = synth.shape_tiling(shape)
st = synth.random_blocks(t)
rng = np.random.default_rng(5)
n1h, n1 = tl.cr_n1_offset(t); n2h, n2 = tl.cr_n2_offset(t); e2h, e2 = tl.cr_e2_offset(t)


class CR:
    pass


cr = CR()
cr.n1_hash, cr.n1 = n1h, rng.uniform(-1, 1, n1) * 0.1
cr.n2_hash, cr.n2 = n2h, rng.uniform(-1, 1, n2) * 0.1
cr.e2_hash, cr.e2 = e2h, rng.uniform(-1, 1, e2) * 0.02
tr = capi.Triples(0)
tr.set_state(st)
tr.set_cr(cr)
for _ in range(2):
    tr.stats(reset=True)
    tr.set_timing(True)
    s4 = tr.run_cr(max_tasks=n)
    s = tr.stats()
    print(shape, n, s4, "fused_ms", s["fused_ms"], "TF", s["flops"] / max(s["fused_ms"], 1e-9) * 1e-9, flush=True)
