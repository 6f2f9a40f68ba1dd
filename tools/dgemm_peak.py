"""cuBLAS DGEMM throughput on this GPU (roofline denominator for the FP64-bound kernels)."""
import json, sys, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"dgemm_n": n, "ms": best, "fp64_tflops": 2.0 * n ** 3 / best * 1e-9}))
