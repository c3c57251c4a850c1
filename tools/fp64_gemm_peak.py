"""cuBLAS DGEMM throughput (burst + sustained) on this B200: second fp64 roofline reference."""
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / (best * 1e-3) / 1e12
t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    c = a @ b; k += 1
    if k % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = 2 * n**3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
print(json.dumps({"dgemm_tflops_burst": round(burst, 2), "dgemm_tflops_sustained": round(sus, 2), "n": n}))
