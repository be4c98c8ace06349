"""CUDA-event timing of lecb_attn_fwd at the ViT shapes (BASELINE configs 3 and 5) and the text-tower shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lecb200 import ops  # noqa: E402

for name, b, t, heads, causal in (("ViT-B/16@448", 128, 785, 12, False), ("ViT-L/14@448", 64, 1025, 16, False),
                                  ("text L=77", 224, 77, 8, True)):
    w = heads * 64
    qkv = torch.randn((b * t, 3 * w), device="cuda").bfloat16()
    out = torch.empty((b * t, w), device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.attn_fwd(qkv, b, t, w, heads, causal=causal, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.attn_fwd(qkv, b, t, w, heads, causal=causal, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * b * heads * t * t * 64 * (0.5 if causal else 1.0)
    print(f"{name}: B={b} T={t} heads={heads} {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (algorithmic)")
