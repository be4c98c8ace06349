"""Tiny launcher for ncu captures of single kernels: python tools/ncu_one.py <what> [pair=0|1]
what: stem | stem_u8 | conv3 (layer3 3x3) | reduce | expand | attn"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lecb200 import _lib, ops  # noqa: E402

what = sys.argv[1]
if len(sys.argv) > 2:
    _lib.lib.lecb_set_pair_gemm(int(sys.argv[2]))
dev = "cuda"
torch.manual_seed(0)
if what in ("stem", "stem_u8"):
    w27 = torch.randn((27, 32), device=dev)
    bias = torch.randn((32,), device=dev)
    if what == "stem":
        img = torch.randn((256, 3, 448, 448), device=dev)
        fn = lambda: ops.stem_conv1(img, w27, bias)
    else:
        img = torch.randint(0, 256, (256, 448, 448, 3), device=dev, dtype=torch.uint8)
        fn = lambda: ops.stem_conv1_u8(img, w27, bias)
elif what == "conv3":
    x = torch.randn((256, 28, 28, 256), device=dev).bfloat16()
    wt = (torch.randn((256, 3, 3, 256), device=dev) * 0.02).bfloat16()
    bias = torch.randn((256,), device=dev)
    fn = lambda: ops.conv3x3(x, wt, bias, relu=True)
elif what in ("reduce", "expand"):
    m, n, k = (200704, 256, 1024) if what == "reduce" else (200704, 1024, 256)
    a = torch.randn((m, k), device=dev).bfloat16()
    w = (torch.randn((n, k), device=dev) * k ** -0.5).bfloat16()
    bias = torch.randn((n,), device=dev)
    res = torch.randn((m, n), device=dev).bfloat16() if what == "expand" else None
    fn = lambda: ops.gemm(a, w, bias, residual=res, relu=True)
else:
    raise SystemExit(what)
for _ in range(4):
    fn()
torch.cuda.synchronize()
