"""A/B of the CTA-pair (cta_group::2) kernel against the single-CTA kernel on the wide layers of the headline workload
(and the ViT GEMMs), alternating the two modes inside one process so both see the same clocks / thermal state."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lecb200 import _lib, ops  # noqa: E402


def timeit(fn, iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def ab(name, fn, flops, rounds=6, iters=10):
    t = {0: [], 1: []}
    for mode in (0, 1):
        _lib.lib.lecb_set_pair_gemm(mode)
        for _ in range(3):
            fn()
    for _ in range(rounds):
        for mode in (0, 1):
            _lib.lib.lecb_set_pair_gemm(mode)
            t[mode].append(timeit(fn, iters))
    _lib.lib.lecb_set_pair_gemm(1)
    m0, m1 = sorted(t[0])[len(t[0]) // 2], sorted(t[1])[len(t[1]) // 2]
    print(json.dumps({"layer": name, "single_ms": round(m0, 4), "pair_ms": round(m1, 4), "pair_speedup": round(m0 / m1, 3),
                      "single_tflops": round(flops / m0 / 1e9, 1), "pair_tflops": round(flops / m1 / 1e9, 1)}), flush=True)


def gemm_case(name, m, n, k, res=False, relu=True, gelu=False):
    a = torch.randn((m, k), device="cuda").bfloat16()
    w = (torch.randn((n, k), device="cuda") * k ** -0.5).bfloat16()
    bias = torch.randn((n,), device="cuda")
    r = torch.randn((m, n), device="cuda").bfloat16() if res else None
    out = torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
    ab(name, lambda: ops.gemm(a, w, bias, residual=r, relu=relu, quick_gelu=gelu, out=out), 2.0 * m * n * k)


def gemm_f32res_case(name, m, n, k):
    a = torch.randn((m, k), device="cuda").bfloat16()
    w = (torch.randn((n, k), device="cuda") * k ** -0.5).bfloat16()
    bias = torch.randn((n,), device="cuda")
    r = torch.randn((m, n), device="cuda")
    out = torch.empty((m, n), device="cuda", dtype=torch.float32)
    ab(name, lambda: ops.gemm_f32res(a, w, bias, r, out=out), 2.0 * m * n * k)


def conv_case(name, b, h, w, cin, cout):
    x = torch.randn((b, h, w, cin), device="cuda").bfloat16()
    wt = (torch.randn((cout, 3, 3, cin), device="cuda") * (9 * cin) ** -0.5).bfloat16()
    bias = torch.randn((cout,), device="cuda")
    out = torch.empty((b, h, w, cout), device="cuda", dtype=torch.bfloat16)
    ab(name, lambda: ops.conv3x3(x, wt, bias, relu=True, out=out), 2.0 * b * h * w * cout * 9 * cin)


conv_case("layer3 3x3 256->256 @28", 256, 28, 28, 256, 256)
gemm_case("layer3 reduce 1024->256", 200704, 256, 1024)
gemm_case("layer3 expand 256->1024 +res", 200704, 1024, 256, res=True)
conv_case("layer4 3x3 512->512 @14", 256, 14, 14, 512, 512)
gemm_case("layer4 expand 512->2048 +res", 50176, 2048, 512, res=True)
gemm_case("layer4 reduce 2048->512", 50176, 512, 2048)
gemm_case("attnpool v_proj 2048->2048", 50176, 2048, 2048, relu=False)
gemm_case("layer2.0 downsample 256->512", 802816, 512, 256, relu=False)
gemm_case("ViT-B qkv 768->2304", 100480, 2304, 768, relu=False)
gemm_case("ViT-B fc 768->3072 gelu", 100480, 3072, 768, relu=False, gelu=True)
gemm_case("ViT-L fc 1024->4096 gelu", 131200, 4096, 1024, relu=False, gelu=True)
gemm_f32res_case("ViT-B out-proj 768->768 f32 res", 100480, 768, 768)
gemm_f32res_case("ViT-B mlp proj 3072->768 f32 res", 100480, 768, 3072)
gemm_f32res_case("ViT-L mlp proj 4096->1024 f32 res", 131200, 1024, 4096)
