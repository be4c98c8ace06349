"""CUDA-event timing of lecb_attn_fwd at the ViT shapes (BASELINE configs 3 and 5) and the text-tower shape, for every
setting of the polynomial-exp2 share (lecb_set_attn_poly: 0 = all on the MUFU, n = every n-th score on the FMA pipe),
alternating the settings inside one process; also the max deviation of each setting's output from the all-MUFU output."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lecb200 import _lib, ops  # noqa: E402

MODES = (0, 4, 3, 2)
for name, b, t, heads, causal in (("ViT-B/16@448", 128, 785, 12, False), ("ViT-L/14@448", 64, 1025, 16, False),
                                  ("text L=77", 224, 77, 8, True)):
    w = heads * 64
    qkv = torch.randn((b * t, 3 * w), device="cuda").bfloat16()
    out = torch.empty((b * t, w), device="cuda", dtype=torch.bfloat16)
    fl = 4.0 * b * heads * t * t * 64 * (0.5 if causal else 1.0)
    times = {m: [] for m in MODES}
    outs = {}
    for m in MODES:
        _lib.lib.lecb_set_attn_poly(m)
        for _ in range(3):
            ops.attn_fwd(qkv, b, t, w, heads, causal=causal, out=out)
        outs[m] = out.float().clone()
    for _ in range(5):
        for m in MODES:
            _lib.lib.lecb_set_attn_poly(m)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.attn_fwd(qkv, b, t, w, heads, causal=causal, out=out)
            e1.record()
            torch.cuda.synchronize()
            times[m].append(e0.elapsed_time(e1) / 10)
    row = {"shape": name, "B": b, "T": t, "heads": heads}
    for m in MODES:
        ms = sorted(times[m])[len(times[m]) // 2]
        row[f"poly{m}"] = {"us": round(ms * 1e3, 1), "tflops": round(fl / ms / 1e9, 1),
                           "max_abs_dev_vs_mufu": float((outs[m] - outs[0]).abs().max())}
    print(json.dumps(row), flush=True)
_lib.lib.lecb_set_attn_poly(0)      # back to the default
