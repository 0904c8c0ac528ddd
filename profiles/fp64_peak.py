"""FP64 GEMM ceiling of this GPU (cuBLAS DGEMM through torch.matmul), the denominator for the DMMA-bound
Schur updates ('measure a cuBLAS DGEMM 8192^3 once on the box')."""
import json
import torch

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    (a @ b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c = a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"dgemm_n": n, "ms": best, "fp64_tflops": 2 * n**3 / best / 1e9}))
