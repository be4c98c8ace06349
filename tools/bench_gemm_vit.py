"""Micro-benchmark of lecb_gemm_bf16 at the ViT-B/16 / text-tower shapes, epilogue variants side by side with torch (cuBLAS)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lecb200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for m, n, k in ((100480, 2304, 768), (100480, 3072, 768), (100480, 768, 3072), (100480, 768, 768), (12320, 1536, 512),
                (12320, 2048, 512), (12320, 512, 2048), (12320, 512, 512), (4928, 2048, 512)):
    a = torch.randn((m, k), device="cuda").bfloat16()
    w = (torch.randn((n, k), device="cuda") * k ** -0.5).bfloat16()
    bias = torch.randn((n,), device="cuda")
    res = torch.randn((m, n), device="cuda")
    fl = 2.0 * m * n * k
    row = [f"M={m:6d} N={n:4d} K={k:4d}"]
    for name, fn in (("plain", lambda: ops.gemm(a, w)), ("bias", lambda: ops.gemm(a, w, bias)),
                     ("gelu", lambda: ops.gemm(a, w, bias, quick_gelu=True)), ("f32res", lambda: ops.gemm_f32res(a, w, bias, res)),
                     ("cublas", lambda: torch.matmul(a, w.t()))):
        ms = timeit(fn)
        row.append(f"{name} {fl / ms / 1e9:6.0f}")
    print("  ".join(row), "TF/s", flush=True)
