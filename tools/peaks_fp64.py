"""GPU box, measurement only (a denominator, never on the product path): FP64 GEMM throughput of cuBLAS through torch.matmul,
quoted beside the dense Cholesky's DMMA utilisation (BASELINE.md section 2 asks for a measured FP64 peak), and an FP64 copy
bandwidth for reference."""
import json
import sys
import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    torch.matmul(a, b)
e1.record(); torch.cuda.synchronize()
sustained = e0.elapsed_time(e1) / 10
print(json.dumps({"fp64_gemm_tflops_burst": 2.0 * n ** 3 / (best * 1e-3) / 1e12, "fp64_gemm_tflops_sustained": 2.0 * n ** 3 / (sustained * 1e-3) / 1e12,
                  "n": n, "how": "torch.matmul float64 %d^3 (cuBLAS DGEMM): best of 5 and 10 back to back, CUDA events" % n,
                  "gpu": torch.cuda.get_device_name(0)}))
