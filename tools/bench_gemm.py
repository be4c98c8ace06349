"""Micro-benchmark of lecb_gemm_bf16 / lecb_conv3x3_bf16 on RN101@448 layer shapes (B images)."""
import sys
import torch
from lecb200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


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


def gemm_case(name, m, n, k, res=False):
    a = torch.randn((m, k), device="cuda").bfloat16()
    w = torch.randn((n, k), device="cuda").bfloat16()
    bias = torch.randn((n,), device="cuda")
    r = torch.randn((m, n), device="cuda").bfloat16() if res else None
    out = torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.gemm(a, w, bias, residual=r, relu=True, out=out))
    fl = 2.0 * m * n * k
    by = 2.0 * (m * k + n * k + m * n * (2 if res else 1))
    ms_t = timeit(lambda: torch.relu_(torch.addmm(bias.bfloat16(), a, w.t())))
    print(f"{name:34s} M={m:8d} N={n:5d} K={k:5d}  {ms:8.3f} ms  {fl/ms/1e9:8.1f} TF/s  {by/ms/1e6:8.1f} GB/s   torch {ms_t:8.3f} ms", flush=True)


def conv_case(name, b, h, c_in, c_out):
    x = torch.randn((b, h, h, c_in), device="cuda").bfloat16()
    w = torch.randn((c_out, 3, 3, c_in), device="cuda").bfloat16()
    bias = torch.randn((c_out,), device="cuda")
    out = torch.empty((b, h, h, c_out), device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.conv3x3(x, w, bias, out=out))
    fl = 2.0 * b * h * h * c_out * 9 * c_in
    by = 2.0 * (b * h * h * (c_in + c_out) + 9 * c_in * c_out)
    xc = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    wc = w.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    ms_t = timeit(lambda: torch.nn.functional.conv2d(xc, wc, bias.bfloat16(), padding=1))
    print(f"{name:34s} B={b} H={h:4d} Cin={c_in:4d} Cout={c_out:4d}  {ms:8.3f} ms  {fl/ms/1e9:8.1f} TF/s  {by/ms/1e6:8.1f} GB/s   cudnn {ms_t:8.3f} ms", flush=True)


print(f"batch {B}")
conv_case("stem conv2 3x3 32->32 @224", B, 224, 32, 32)
conv_case("stem conv3 3x3 32->64 @224", B, 224, 32, 64)
gemm_case("layer1 1x1 64->64", B * 112 * 112, 64, 64)
gemm_case("layer1 1x1 64->256 +res", B * 112 * 112, 256, 64, True)
gemm_case("layer1 1x1 256->64", B * 112 * 112, 64, 256)
conv_case("layer1 3x3 64 @112", B, 112, 64, 64)
gemm_case("layer2 1x1 512->128", B * 56 * 56, 128, 512)
gemm_case("layer2 1x1 128->512 +res", B * 56 * 56, 512, 128, True)
conv_case("layer2 3x3 128 @56", B, 56, 128, 128)
gemm_case("layer3 1x1 1024->256", B * 28 * 28, 256, 1024)
gemm_case("layer3 1x1 256->1024 +res", B * 28 * 28, 1024, 256, True)
conv_case("layer3 3x3 256 @28", B, 28, 256, 256)
gemm_case("layer4 1x1 2048->512", B * 14 * 14, 512, 2048)
gemm_case("layer4 1x1 512->2048 +res", B * 14 * 14, 2048, 512, True)
conv_case("layer4 3x3 512 @14", B, 14, 512, 512)
gemm_case("v_proj 2048->2048", B * 196, 2048, 2048)
gemm_case("c_proj 2048->512", B * 196, 512, 2048)
gemm_case("square 8192", 8192, 8192, 8192)
