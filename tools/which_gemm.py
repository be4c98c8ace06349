"""Which kernel does lecb_gemm_bf16 route a shape to?  (torch.profiler / CUPTI kernel names + CUDA-event time, pair on / off.)
    python tools/which_gemm.py M N K [M N K ...]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lecb200 import _lib, ops  # noqa: E402

args = [int(v) for v in sys.argv[1:]] or [12320, 1536, 512, 12320, 2048, 512, 12320, 512, 2048, 1792, 1536, 512]
for i in range(0, len(args), 3):
    m, n, k = args[i:i + 3]
    a = torch.randn((m, k), device="cuda").bfloat16()
    w = torch.randn((n, k), device="cuda").bfloat16()
    for mode in (1, 0):
        _lib.lib.lecb_set_pair_gemm(mode)
        for _ in range(3):
            ops.gemm(a, w)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            ops.gemm(a, w)
            torch.cuda.synchronize()
        names = [e.key[:90] for e in prof.key_averages() if "gemm" in e.key]
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            ops.gemm(a, w)
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) * 50
        print(f"M={m} N={n} K={k} pair_switch={mode}: {us:.1f} us  {2e-6 * m * n * k / us:.0f} TF/s  {names}", flush=True)
_lib.lib.lecb_set_pair_gemm(1)
