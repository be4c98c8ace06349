"""HBM-roofline evidence for the vectorised row kernels at scaled inputs (real shapes such as the
[512,80] ASL batch are launch-latency-bound, so the bandwidth claim is demonstrated on >= 0.5 GB of traffic).
Prints one JSON line per kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lecb200 import ops  # noqa: E402

PEAK = 6540.2
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = json.load(open(p))["hbm_gbs"]


ITERS_CAP = int(os.environ.get("LECB_ROWOPS_ITERS", "0"))      # > 0: few launches per kernel (the ncu pass)


def timeit(fn, iters=20):
    if ITERS_CAP:
        iters = min(iters, ITERS_CAP)
    for _ in range(1 if ITERS_CAP else 3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def report(name, nbytes, ms, shape):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "shape": shape, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
                      "achieved_GBs": round(gbs, 1), "peak_GBs": PEAK, "frac": round(gbs / PEAK, 3)}), flush=True)


dev = "cuda"
# head aggregation: B images x P patches x (3K dots in, 2K maps out) + logits
B, P, K = 4096, 196, 80
dots = torch.randn((B * P, 3 * K), device=dev) * 0.1
ssq = torch.rand((B * P,), device=dev) + 0.5
ms = timeit(lambda: ops.head_aggregate(dots, B, P, K, 3, row_sumsq=ssq, want_maps=True))
report("head_aggregate (evidence, maps)", B * P * (3 * K + 2 * K) * 4 + B * P * 4 + B * K * 4, ms, f"B={B} P={P} K={K}")
ms = timeit(lambda: ops.head_aggregate(dots, B, P, K, 3, row_sumsq=ssq, want_maps=False))
# without the map outputs the positive-prompt columns are never read: 2K of the 3K floats per row
report("head_aggregate (evidence, no maps)", B * P * 2 * K * 4 + B * P * 4 + B * K * 4, ms, f"B={B} P={P} K={K}")
del dots
# ASL fwd+bwd
n = 1 << 21
x = torch.randn((n, 80), device=dev) * 2
y = (torch.rand((n, 80), device=dev) < 0.04).float()
ms = timeit(lambda: ops.asl_fwd_bwd(x, y))
report("asl_fwd_bwd", n * 80 * 12, ms, f"[{n},80]")
ms = timeit(lambda: ops.ranking_fwd_bwd(x[: 1 << 19], y[: 1 << 19], 1.0, 1.0))
report("ranking_fwd_bwd", (1 << 19) * 80 * 12, ms, f"[{1 << 19},80]")
del x, y
# L2 norm
rows, d = 196 * 4096, 512
xb = torch.randn((rows, d), device=dev).bfloat16()
ms = timeit(lambda: ops.l2norm_rows(xb))
report("l2norm_rows bf16->bf16", rows * d * 4, ms, f"[{rows},{d}]")
xf = torch.randn((rows // 2, d), device=dev)
ms = timeit(lambda: ops.l2norm_rows(xf))
report("l2norm_rows f32->f32", rows // 2 * d * 8, ms, f"[{rows // 2},{d}]")
# LayerNorm fwd (fp32 in, bf16 out)
g, b_ = torch.ones(d, device=dev), torch.zeros(d, device=dev)
ms = timeit(lambda: ops.layernorm(xf, g, b_))
report("layernorm_fwd f32->bf16", rows // 2 * d * 6, ms, f"[{rows // 2},{d}]")
del xf
# avgpool
xp = torch.randn((256, 112, 112, 256), device=dev).bfloat16()
ms = timeit(lambda: ops.avgpool2x2(xp))
report("avgpool2x2", xp.numel() * 2 * 1.25, ms, "B=256 112x112x256")
del xp
# quick gelu
v = torch.randn((1 << 28,), device=dev).bfloat16()
ms = timeit(lambda: ops.quick_gelu_fwd(v))
report("quick_gelu_fwd", v.numel() * 4, ms, f"[{v.numel()}]")
# stem conv1
img = torch.randn((256, 3, 448, 448), device=dev)
w27 = torch.randn((27, 32), device=dev)
bias = torch.randn((32,), device=dev)
ms = timeit(lambda: ops.stem_conv1(img, w27, bias))
report("stem_conv1", img.numel() * 4 + 256 * 224 * 224 * 32 * 2, ms, "B=256 448x448")
u8 = torch.randint(0, 256, (256, 448, 448, 3), device=dev, dtype=torch.uint8)
ms = timeit(lambda: ops.stem_conv1_u8(u8, w27, bias))
report("stem_conv1_u8", u8.numel() + 256 * 224 * 224 * 32 * 2, ms, "B=256 448x448 uint8 NHWC")
del img, u8
# KL consistency term and the co-occurrence ranking loss at scaled inputs
n = 1 << 19
x = torch.randn((n, 80), device=dev) * 2
xm = x + torch.randn_like(x) * 0.3
y = (torch.rand((n, 80), device=dev) < 0.04).float()
wt = torch.rand((80, 80), device=dev) + 0.5
ms = timeit(lambda: ops.kl_softmax_fwd_bwd(x, xm, 1.0))
report("kl_softmax_fwd_bwd", n * 80 * 12, ms, f"[{n},80]")
ms = timeit(lambda: ops.ranking_cooc_fwd_bwd(x, y, wt, 1.0, 1.0))
report("ranking_cooc_fwd_bwd", n * 80 * 12, ms, f"[{n},80]")
del x, xm, y
# test-time windows: 116 windows (scales 2, 3 + the whole image) of a 375 x 500 image -> 448 x 448 uint8
import numpy as np  # noqa: E402
from lecb200 import windows as WN  # noqa: E402
h, w = 375, 500
wins = [WN.whole_image(h, w)] + [q for s in (2, 3) for q in WN.sliding_windows(h, w, s)]
img8 = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(h, w, 3), dtype=np.uint8)).cuda()
ms = timeit(lambda: WN.crop_resize(img8, wins, 448), iters=10)
report("crop_resize_u8 (plan on host + 2 kernels)", len(wins) * 448 * 448 * 3, ms, f"{len(wins)} windows of {h}x{w} -> 448x448")
