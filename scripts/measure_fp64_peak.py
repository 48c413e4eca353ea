"""Measures the FP64 GEMM throughput of the GPU (cuBLAS DGEMM through torch.matmul): the denominator for the
tensor-path fraction of sigma_ab_kernel (MEASURED_PEAKS.json has no FP64 entry).  Prints one JSON line."""
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    (a @ b)
torch.cuda.synchronize()
best = 0.0
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
t0 = time.time(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); k = 0
while time.time() - t0 < 3.0:
    c = a @ b; k += 1
e1.record(); torch.cuda.synchronize()
sustained = k * 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
print(json.dumps({"fp64_gemm_tflops_burst": best, "fp64_gemm_tflops_sustained": sustained, "n": n, "how": "torch.matmul float64 8192^3"}))
