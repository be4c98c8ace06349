"""tcgen05 GEMM / implicit-GEMM conv parity on the B200 (through the C ABI).

Reference: fp32 torch matmul / conv2d on the same bf16-rounded operands; the kernel accumulates in
fp32 so the only difference is summation order and the final bf16 rounding of the output:
tolerance = 2^-8 relative to the output scale (one bf16 ulp) + small absolute slack."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def _check(got, want, what):
    got = got.float()
    scale = want.abs().max().item() + 1e-6
    err = (got - want).abs().max().item()
    assert err <= scale * (2.0 ** -7), f"{what}: max err {err:.4g} vs scale {scale:.4g}"


GEMM_SHAPES = [
    # (M, N, K) : tile-aligned, ragged M, small N, N tail (240), K = 32-multiple, multi-wave
    (128, 64, 64), (256, 128, 128), (200, 256, 512), (1000, 512, 2048), (333, 240, 512),
    (128, 32, 96), (4096, 2048, 256), (777, 1024, 1024), (50176, 64, 256), (19000, 256, 64), (64, 8, 32),
]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_plain(m, n, k):
    from lecb200 import ops
    a = _rand((m, k), 1).bfloat16()
    w = _rand((n, k), 2, k ** -0.5).bfloat16()
    out = ops.gemm(a, w)
    torch.cuda.synchronize()
    _check(out, a.float() @ w.float().t(), f"gemm {m}x{n}x{k}")


@pytest.mark.parametrize("m,n,k", [(300, 256, 512), (1111, 64, 64), (515, 2048, 2048)])
def test_gemm_epilogues(m, n, k):
    from lecb200 import ops
    a = _rand((m, k), 3).bfloat16()
    w = _rand((n, k), 4, k ** -0.5).bfloat16()
    bias = _rand((n,), 5)
    res = _rand((m, n), 6).bfloat16()
    base = a.float() @ w.float().t() + bias
    _check(ops.gemm(a, w, bias), base, "bias")
    _check(ops.gemm(a, w, bias, relu=True), base.relu(), "bias+relu")
    _check(ops.gemm(a, w, bias, residual=res, relu=True), (base + res.float()).relu(), "bias+res+relu")
    _check(ops.gemm(a, w, bias, quick_gelu=True), base * torch.sigmoid(1.702 * base), "quickgelu")
    f32 = ops.gemm(a, w, bias, out_f32=True)
    assert f32.dtype == torch.float32
    assert (f32 - base).abs().max().item() <= 1e-3 * (base.abs().max().item() + 1)
    ssq = torch.zeros((m,), device="cuda")
    out = ops.gemm(a, w, bias, row_sumsq=ssq)
    torch.cuda.synchronize()
    want = out.float().pow(2).sum(-1)
    assert ((ssq - want).abs() / (want + 1e-6)).max().item() < 1e-4


CONV_SHAPES = [
    # (B, H, W, Cin, Cout)
    (2, 8, 8, 64, 64), (1, 14, 14, 512, 512), (3, 28, 28, 256, 256), (2, 56, 56, 128, 128),
    (2, 32, 32, 32, 32), (1, 64, 64, 32, 64), (5, 7, 7, 512, 512), (2, 112, 112, 64, 64), (1, 5, 9, 64, 128),
]


@pytest.mark.parametrize("b,h,w,cin,cout", CONV_SHAPES)
def test_conv3x3(b, h, w, cin, cout):
    from lecb200 import ops
    x = _rand((b, h, w, cin), 7).bfloat16()
    wt = _rand((cout, 3, 3, cin), 8, (9 * cin) ** -0.5).bfloat16()
    bias = _rand((cout,), 9, 0.1)
    out = ops.conv3x3(x, wt, bias, relu=True)
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, padding=1)
    want = want.relu().permute(0, 2, 3, 1)
    _check(out, want, f"conv {b}x{h}x{w}x{cin}->{cout}")


def test_gemm_rejects_bad_args():
    from lecb200 import LecbError, ops
    a = torch.zeros((16, 40), device="cuda", dtype=torch.bfloat16)
    w = torch.zeros((8, 40), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(LecbError):
        ops.gemm(a, w)          # K not a multiple of 32


@pytest.mark.parametrize("m,n,k", [(300, 256, 512), (1111, 64, 64), (515, 768, 768), (50176, 240, 512), (100, 32, 64),
                                   (4000, 1024, 3072)])
def test_gemm_f32_out_staged(m, n, k):
    """fp32 output (+ fp32 residual) through the staged TMA-store epilogue (32-column blocks), N tails, ragged M."""
    from lecb200 import ops
    a = _rand((m, k), 11).bfloat16()
    w = _rand((n, k), 12, k ** -0.5).bfloat16()
    bias = _rand((n,), 13)
    res = _rand((m, n), 14)
    base = a.float() @ w.float().t() + bias
    tol = 2e-3 * (base.abs().max().item() + 1)
    out = ops.gemm(a, w, bias, out_f32=True)
    assert (out - base).abs().max().item() <= tol
    out = ops.gemm_f32res(a, w, bias, res)
    assert (out - (base + res)).abs().max().item() <= tol
    ssq = torch.zeros((m,), device="cuda")
    out = ops.gemm(a, w, bias, out_f32=True, row_sumsq=ssq)
    torch.cuda.synchronize()
    want = out.pow(2).sum(-1)
    assert ((ssq - want).abs() / (want + 1e-6)).max().item() < 1e-4
    g = ops.gemm(a, w, bias, quick_gelu=True, out_f32=True)
    assert (g - base * torch.sigmoid(1.702 * base)).abs().max().item() <= tol


@pytest.mark.parametrize("m,n,k", [(100003, 1024, 256), (90000, 512, 128), (80000, 256, 64), (77000, 768, 128),
                                   (76000, 64, 256), (100352, 240, 512)])
def test_gemm_resident_weights_mode(m, n, k):
    """Tall problems (>= 4 m tiles per SM) with small weights run with the CTA's W tile resident in shared memory and a
    CTA-owns-one-n-tile schedule (1, 2, 3 and 4 n tiles, ragged M, N tail, bf16 + fp32 epilogues)."""
    from lecb200 import ops
    a = _rand((m, k), 21).bfloat16()
    w = _rand((n, k), 22, k ** -0.5).bfloat16()
    bias = _rand((n,), 23)
    res = _rand((m, n), 24).bfloat16()
    base = a.float() @ w.float().t() + bias
    _check(ops.gemm(a, w, bias, residual=res, relu=True), (base + res.float()).relu(), "resident bias+res+relu")
    _check(ops.gemm(a, w, bias), base, "resident bias")
    f32 = ops.gemm(a, w, bias, out_f32=True)
    assert (f32 - base).abs().max().item() <= 2e-3 * (base.abs().max().item() + 1)


def test_conv3x3_resident_weights_mode():
    from lecb200 import ops
    b, h, w, cin, cout = 8, 112, 112, 64, 64
    x = _rand((b, h, w, cin), 31).bfloat16()
    wt = _rand((cout, 3, 3, cin), 32, (9 * cin) ** -0.5).bfloat16()
    bias = _rand((cout,), 33, 0.1)
    out = ops.conv3x3(x, wt, bias, relu=True)
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, padding=1)
    _check(out, want.relu().permute(0, 2, 3, 1), "resident conv")


@pytest.mark.parametrize("b,h,w", [(2, 64, 64), (3, 224, 224), (1, 450, 446), (5, 32, 96)])
def test_stem_conv1_tensor_core(b, h, w):
    """Stem conv1 (3x3 / stride 2 / pad 1, 3 -> 32, folded BN + ReLU) as a tcgen05 implicit GEMM vs torch conv2d on the
    same bf16-rounded operands, and vs the CUDA-core kernel (ragged last tile, odd image counts, non-square images)."""
    import os
    from lecb200 import ops
    x = _rand((b, 3, h, w), 41)
    wt = _rand((32, 3, 3, 3), 42, 27 ** -0.5)
    bias = _rand((32,), 43, 0.1)
    w27 = wt.permute(1, 2, 3, 0).reshape(27, 32).contiguous()
    out = ops.stem_conv1(x, w27, bias)
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(x.bfloat16().float(), wt.bfloat16().float(), bias, stride=2, padding=1).relu().permute(0, 2, 3, 1)
    _check(out, want, f"stem conv1 tc {b}x{h}x{w}")
    os.environ["LECB_STEM_CUDA_CORES"] = "1"
    try:
        ref = ops.stem_conv1(x, w27, bias)
    finally:
        del os.environ["LECB_STEM_CUDA_CORES"]
    torch.cuda.synchronize()
    # the CUDA-core kernel keeps fp32 inputs/weights: differences are the bf16 rounding of the operands
    assert (out.float() - ref.float()).abs().max().item() <= 3e-2 * (ref.float().abs().max().item() + 1e-6)


@pytest.mark.parametrize("b,h,w,cout", [(8, 112, 112, 64), (4, 224, 112, 64), (24, 56, 56, 64), (20, 50, 64, 64), (6, 112, 112, 40),
                                        (4, 224, 112, 128), (8, 112, 112, 96)])
def test_conv3x3_halo_tile_mode(b, h, w, cout):
    """64-channel 3x3 convs run in halo-tile mode (patch tiles, three shifted halo copies per stage, TMA zero-fill as the
    conv padding, 4-D TMA store that clips ragged patches): exact tilings, ragged H, 8x16 and 16x8 patches, Cout < 64."""
    from lecb200 import ops
    x = _rand((b, h, w, 64), 51).bfloat16()
    wt = _rand((cout, 3, 3, 64), 52, (9 * 64) ** -0.5).bfloat16()
    bias = _rand((cout,), 53, 0.1)
    out = ops.conv3x3(x, wt, bias, relu=True)
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, padding=1)
    _check(out, want.relu().permute(0, 2, 3, 1), f"halo conv {b}x{h}x{w}->{cout}")


@pytest.mark.parametrize("b,h,w,cout", [(4, 224, 224, 32), (4, 224, 224, 64), (9, 112, 96, 64), (16, 50, 64, 32), (6, 112, 112, 40),
                                        (3, 224, 224, 24)])
def test_conv3x3_halo_tile_mode_cin32(b, h, w, cout):
    """The Ci = 32 stem convs (T:385-399 via M:144-151) in halo-tile mode with 64-byte rows (64B swizzle, K step 32,
    two MMAs per tap): exact tilings, ragged H, both patch shapes, Cout tails."""
    from lecb200 import ops
    x = _rand((b, h, w, 32), 61).bfloat16()
    wt = _rand((cout, 3, 3, 32), 62, (9 * 32) ** -0.5).bfloat16()
    bias = _rand((cout,), 63, 0.1)
    out = ops.conv3x3(x, wt, bias, relu=True)
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, padding=1)
    _check(out, want.relu().permute(0, 2, 3, 1), f"halo conv cin32 {b}x{h}x{w}->{cout}")


@pytest.mark.parametrize("b,h,w,cin,cout", [(4, 224, 224, 32, 64), (9, 112, 96, 32, 32), (8, 112, 112, 64, 64), (4, 224, 112, 64, 128),
                                            (20, 50, 64, 64, 64), (6, 112, 112, 32, 40), (2, 8, 8, 64, 64), (2, 56, 56, 128, 128)])
def test_conv3x3_fused_avgpool(b, h, w, cin, cout):
    """conv3x3 + ReLU + 2x2 average pool (M:147 stem avgpool, M:27/46 Bottleneck avgpool): fused epilogue in halo-tile
    mode (both patch shapes, ragged H, Cout tail, two-block tiles), two kernels elsewhere (last two shapes); the pooled
    fp32 reference is rounded to bf16 once."""
    from lecb200 import _lib, ops
    x = _rand((b, h, w, cin), 71).bfloat16()
    wt = _rand((cout, 3, 3, cin), 72, (9 * cin) ** -0.5).bfloat16()
    bias = _rand((cout,), 73, 0.1)
    out = ops.conv3x3(x, wt, bias, relu=True, pool=True)
    torch.cuda.synchronize()
    assert tuple(out.shape) == (b, h // 2, w // 2, cout)
    fus = _lib.lib.lecb_conv3x3_pool_fusable(b, h, w, cin, cout)
    assert fus == (1 if (cin <= 64 and b * h * w >= 128 * 2 * 148) else 0)
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, padding=1).relu()
    want = torch.nn.functional.avg_pool2d(want, 2).permute(0, 2, 3, 1)
    _check(out, want, f"conv+pool {b}x{h}x{w} {cin}->{cout}")


@pytest.mark.parametrize("b,h,w,cin,cout", [(8, 112, 112, 128, 128), (27, 56, 56, 128, 128), (8, 112, 112, 128, 96), (7, 112, 112, 256, 128)])
def test_conv3x3_paired_m_tiles(b, h, w, cin, cout):
    """128-wide im2col convs with an even number (>= 4 per SM) of m tiles run in paired-tile mode (two A tiles and one W
    tile per stage, four accumulators): exact and ragged M, Cout tail, Cin 128 / 256."""
    from lecb200 import ops
    x = _rand((b, h, w, cin), 81).bfloat16()
    wt = _rand((cout, 3, 3, cin), 82, (9 * cin) ** -0.5).bfloat16()
    bias = _rand((cout,), 83, 0.1)
    out = ops.conv3x3(x, wt, bias, relu=True)
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, padding=1)
    _check(out, want.relu().permute(0, 2, 3, 1), f"paired conv {b}x{h}x{w} {cin}->{cout}")
