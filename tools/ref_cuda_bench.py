"""The reference's own CUDA kernels (sd_t_total.cu + memory.cu, compiled unmodified into oracle/_ref/libsd_t_ref.so)
timed beside libnwc_triples.so on the same B200, through the SAME call sequence a Fortran rank makes for one tuple
(ccsd_t_gpu.F:135-220): initmemmodule, dev_mem_s/d, 9 x sd_t_s1, 9 x sd_t_d1, 9 x sd_t_d2, compute_en, dev_release.
Host buffers, so every call pays its H2D copy in both libraries.  Tile edge T <= 32 (the reference's singles kernel
overflows shared memory above that).   python tools/ref_cuda_bench.py [T] [reps]"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nwchem_b200 import capi
T = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PD = C.POINTER(C.c_double)
pd = lambda a: a.ctypes.data_as(PD)
rng = np.random.default_rng(7)
t1 = rng.standard_normal(T * T); v4 = rng.standard_normal(T ** 4) * 0.1; t4 = rng.standard_normal(T ** 4) * 0.1
eps = [np.sort(rng.uniform(-2.0, -0.4, T)) for _ in range(3)] + [np.sort(rng.uniform(0.1, 3.0, T)) for _ in range(3)]
d = [C.c_long(T) for _ in range(7)]
r = [C.byref(x) for x in d]
flops = 9 * 2.0 * T ** 6 + 18 * 2.0 * T ** 7


def one_tuple(lib):
    e = np.zeros(2); dummy = np.zeros(1)
    lib.initmemmodule_()
    lib.dev_mem_s_(*r[:6]); lib.dev_mem_d_(*r[:6])
    for k in range(1, 10):
        getattr(lib, f"sd_t_s1_{k}_cuda_")(*r[:6], None, pd(t1), pd(v4))
    for k in range(1, 10):
        getattr(lib, f"sd_t_d1_{k}_cuda_")(*r[:7], None, pd(t4), pd(v4))
    for k in range(1, 10):
        getattr(lib, f"sd_t_d2_{k}_cuda_")(*r[:7], None, pd(t4), pd(v4))
    f = C.c_double(1.0)
    lib.compute_en_(C.byref(f), pd(e), *[pd(x) for x in eps], *r[:6], pd(dummy), pd(dummy))
    lib.dev_release_(); lib.finalizememmodule_()
    return e


out = {"tile": T, "flops_per_tuple": flops, "calls": "9 s1 + 9 d1 + 9 d2 + compute_en, host operands"}
libs = [("libnwc_triples", capi.lib())]
ref_path = os.path.join(root, "oracle", "_ref", "libsd_t_ref.so")
if os.path.exists(ref_path):
    libs.append(("reference_sd_t_total_cu", C.CDLL(ref_path)))
energies = {}
for name, lib in libs:
    one_tuple(lib)   # warm-up (context, allocations)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); e = one_tuple(lib); best = min(best, time.perf_counter() - t0)
    energies[name] = e
    out[name] = {"seconds_per_tuple": best, "tflops": flops / best * 1e-12, "e1": float(e[0]), "e2": float(e[1])}
if len(libs) == 2:
    a, b = energies["libnwc_triples"], energies["reference_sd_t_total_cu"]
    out["relative_energy_difference"] = [float(abs(a[i] - b[i]) / abs(b[i])) for i in range(2)]
    out["speedup"] = out["reference_sd_t_total_cu"]["seconds_per_tuple"] / out["libnwc_triples"]["seconds_per_tuple"]
print(json.dumps(out))
