"""cuBLAS DGEMM / copy roofline probes (torch is plumbing here): prints one JSON line."""
import json, torch
def best(f, n=5):
    f(); torch.cuda.synchronize(); b = 1e30
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); e1.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
out = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = best(lambda: torch.matmul(a, b)); out[f"cublas_dgemm_{n}_tflops"] = round(2 * n**3 / ms * 1e-9, 2)
# sustained 3 s
n = 8192
import time
t0 = time.time(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); k = 0
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(4): torch.matmul(a, b); k += 1
    torch.cuda.synchronize()
e1.record(); e1.synchronize()
out["cublas_dgemm_8192_tflops_sustained"] = round(2 * n**3 * k / e0.elapsed_time(e1) * 1e-9, 2)
print(json.dumps(out))
