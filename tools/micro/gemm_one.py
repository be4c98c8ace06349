"""One GEMM shape, a few launches (target for an ncu --set full capture): python gemm_one.py M N K [res]"""
import sys
import torch
from lecb200 import ops

m, n, k = [int(v) for v in sys.argv[1:4]]
res = len(sys.argv) > 4 and sys.argv[4] == "res"
a = torch.randn((m, k), device="cuda").bfloat16()
w = (torch.randn((n, k), device="cuda") * k ** -0.5).bfloat16()
bias = torch.randn((n,), device="cuda")
r = torch.randn((m, n), device="cuda").bfloat16() if res else None
out = torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.gemm(a, w, bias, residual=r, relu=True, out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    ops.gemm(a, w, bias, residual=r, relu=True, out=out)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
by = 2.0 * (m * k + n * k + m * n * (2 if res else 1))
print(f"gemm M={m} N={n} K={k} res={int(res)}: {ms:.4f} ms  {2.0 * m * n * k / ms / 1e9:.0f} TF/s  {by / ms / 1e6:.0f} GB/s")
