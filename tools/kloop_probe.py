"""Asymptotic K-loop efficiency of the fused kernel: one tuple (all ranges R), ONE contraction with a huge K."""
import sys, os, ctypes as C, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi
R = int(sys.argv[1]) if len(sys.argv) > 1 else 40
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
fam_k = [(2, 1)] if len(sys.argv) <= 3 else [(int(x.split(":")[0]), int(x.split(":")[1])) for x in sys.argv[3].split(",")]
l = capi.lib()
rng = np.random.default_rng(0)
ts = rng.standard_normal(K * R ** 3); vs = rng.standard_normal(K * R ** 3)
PD = C.POINTER(C.c_double); pd = lambda a: a.ctypes.data_as(PD)
T = [C.c_long(R) for _ in range(6)]; Kc = C.c_long(K)
eps = [np.sort(rng.uniform(-2, -0.4, R)) for _ in range(3)] + [np.sort(rng.uniform(0.1, 3, R)) for _ in range(3)]
capi.compat_set_timing(True)
for it in range(3):
    capi.compat_stats(reset=True)
    l.initmemmodule_(); l.dev_mem_s_(*[C.byref(x) for x in T]); l.dev_mem_d_(*[C.byref(x) for x in T])
    for fam, k in fam_k:
        fn = getattr(l, f"sd_t_{'d1' if fam == 1 else 'd2'}_{k}_cuda_")
        r = [C.byref(x) for x in T]
        if fam == 1: fn(r[0], r[1], r[2], C.byref(Kc), r[3], r[4], r[5], None, pd(ts), pd(vs))
        else: fn(r[0], r[1], r[2], r[3], r[4], r[5], C.byref(Kc), None, pd(ts), pd(vs))
    e = np.zeros(2); f = C.c_double(1.0)
    l.compute_en_(C.byref(f), pd(e), *[pd(x) for x in eps], *[C.byref(x) for x in T], None, None)
    l.dev_release_(); l.finalizememmodule_()
    s = capi.compat_stats()
    print(json.dumps(dict(R=R, K=K, kernels=fam_k, fused_ms=s["fused_ms"], tflops=s["flops"] / s["fused_ms"] * 1e-9, e1=e[0])), flush=True)
