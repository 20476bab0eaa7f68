"""Library comparator for the assembly kernel: cuBLAS FP64 GEMM V^T V on an explicitly materialised
design matrix V (N x n, n = r p r = 64, N = 2^20), which is what a direct port of als.py:44-55 to cupy/cuBLAS executes
per micro-step — without the cost of forming V.  Prints the time per product next to the algorithmic figures used by
bench.py's roofline (n(n+1)+2n flops per sample)."""
import json

import torch

N, n = 1 << 20, 64
V = torch.randn(N, n, dtype=torch.float64, device="cuda")
y = torch.randn(N, dtype=torch.float64, device="cuda")
Vt = V.t().contiguous()                     # [n][N]: sample-contiguous, the layout of this repo
out = {}
for name, fn in (("V^T V, V row-major (N x n)", lambda: (V.t() @ V, V.t() @ y)),
                 ("V^T V, V sample-contiguous (n x N)", lambda: (Vt @ Vt.t(), Vt @ y))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    out[name] = {"us_per_product": us, "algorithmic_tflops": N * (n * (n + 1) + 2 * n) / (us * 1e-6) / 1e12,
                 "executed_tflops": N * (2 * n * n + 2 * n) / (us * 1e-6) / 1e12, "bytes_read": V.numel() * 8}
print(json.dumps(out))
