"""tcgen05 attention kernel parity on the B200 (through the C ABI, lecb_attn_fwd).

Reference: fp32 torch softmax(QK^T/8)V on the same bf16-rounded q/k/v.  The kernel rounds the
probabilities to bf16 before the PV product (fp32 accumulation) and the output to bf16: tolerance =
2^-7 of the output scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, b, t, w, heads, causal):
    q, k, v = qkv.float().view(b, t, 3, heads, 64).permute(2, 0, 3, 1, 4)          # [b,h,t,64]
    s = (q @ k.transpose(-1, -2)) / 8.0
    if causal:
        s = s + torch.full((t, t), float("-inf"), device=s.device).triu(1)
    return (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(b * t, w)


@pytest.mark.parametrize("b,t,heads,causal", [
    (2, 128, 2, False), (3, 785, 12, False), (2, 1025, 16, False), (1, 50, 1, False), (2, 197, 12, False),
    (5, 77, 8, True), (2, 300, 4, True), (1, 128, 1, True), (2, 257, 2, False),
])
def test_attn_fwd(b, t, heads, causal):
    from lecb200 import ops
    w = heads * 64
    g = torch.Generator(device="cpu").manual_seed(b * 1000 + t)
    qkv = (torch.randn((b * t, 3 * w), generator=g) * 1.5).cuda().bfloat16()
    out = ops.attn_fwd(qkv, b, t, w, heads, causal=causal)
    torch.cuda.synchronize()
    want = _ref(qkv, b, t, w, heads, causal)
    err = (out.float() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= scale * 2.0 ** -7 + 1e-3, f"attn b={b} t={t} h={heads} causal={causal}: err {err:.4g} scale {scale:.4g}"


def test_attn_fwd_peaked_and_partial_rows():
    """Large-magnitude scores (near one-hot softmax) and q_rows < T (class-token-only mode leaves other rows alone)."""
    from lecb200 import ops
    b, t, heads = 2, 785, 12
    w = heads * 64
    g = torch.Generator(device="cpu").manual_seed(7)
    qkv = (torch.randn((b * t, 3 * w), generator=g) * 4.0).cuda().bfloat16()
    want = _ref(qkv, b, t, w, heads, False)
    out = ops.attn_fwd(qkv, b, t, w, heads)
    assert (out.float() - want).abs().max().item() <= want.abs().max().item() * 2.0 ** -6
    sentinel = torch.full((b * t, w), 7.0, device="cuda", dtype=torch.bfloat16)
    out1 = ops.attn_fwd(qkv, b, t, w, heads, q_rows=1, out=sentinel.clone())
    torch.cuda.synchronize()
    o3 = out1.view(b, t, w)
    assert (o3[:, 0].float() - want.view(b, t, w)[:, 0]).abs().max().item() <= want.abs().max().item() * 2.0 ** -6
    assert (o3[:, 1:] == 7.0).all()


@pytest.mark.parametrize("causal", [False, True])
def test_attn_fwd_growing_scores_triggers_rescale(causal):
    """Keys late in the sequence carry much larger scores than the first block: exercises the lazy-rescale path
    (exact re-max of the block + in-place rescale of O in TMEM)."""
    from lecb200 import ops
    b, t, heads = 2, 640, 2
    w = heads * 64
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn((b, t, 3, heads, 64), generator=g)
    ramp = torch.linspace(0.2, 9.0, t).view(1, t, 1, 1)
    x[:, :, 1] *= ramp                                  # key magnitude grows along the sequence
    x[:, :, 0] *= 3.0
    qkv = x.reshape(b * t, 3 * w).cuda().bfloat16()
    out = ops.attn_fwd(qkv, b, t, w, heads, causal=causal)
    torch.cuda.synchronize()
    want = _ref(qkv, b, t, w, heads, causal)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - want).abs().max().item()
    assert err <= want.abs().max().item() * 2.0 ** -6 + 1e-3, f"err {err}"


def test_causal_tc_matches_smem_kernel():
    """Two independent implementations of the text tower's causal attention agree (tcgen05 vs CUDA-core smem kernel)."""
    from lecb200 import ops
    n, l, heads = 37, 77, 8
    w = heads * 64
    g = torch.Generator(device="cpu").manual_seed(3)
    qkv = torch.randn((n * l, 3 * w), generator=g).cuda().bfloat16()
    a = ops.causal_attn(qkv, n, l, w, heads).float()
    b = ops.causal_attn_smem(qkv, n, l, w, heads).float()
    torch.cuda.synchronize()
    assert (a - b).abs().max().item() <= 2.0 ** -6 * b.abs().max().item()


@pytest.mark.parametrize("n,l,heads", [(3, 77, 8), (2, 128, 2), (5, 33, 1), (2, 16, 12), (4, 96, 4)])
def test_attn_causal_bwd(n, l, heads):
    """tcgen05 causal-attention backward vs torch autograd (fp32) on the same bf16-rounded q/k/v/dO, and vs the
    CUDA-core kernel.  P and dS are rounded to bf16 for the MMAs: tolerance 2^-6 of each gradient's scale."""
    from lecb200 import ops
    w = heads * 64
    g = torch.Generator(device="cpu").manual_seed(n * 100 + l)
    qkv = torch.randn((n * l, 3 * w), generator=g).cuda().bfloat16()
    dout = torch.randn((n * l, w), generator=g).cuda().bfloat16()
    x = qkv.float().clone().requires_grad_(True)
    q, k, v = x.view(n, l, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / 8.0 + torch.full((l, l), float("-inf"), device="cuda").triu(1)
    o = (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(n * l, w)
    o.backward(dout.float())
    want = x.grad
    got = ops.causal_attn_bwd(qkv, dout, n, l, w, heads).float()
    ref2 = ops.causal_attn_bwd_smem(qkv, dout, n, l, w, heads).float() if l <= 96 else None
    torch.cuda.synchronize()
    for name, sl in (("dq", slice(0, w)), ("dk", slice(w, 2 * w)), ("dv", slice(2 * w, 3 * w))):
        scale = want[:, sl].abs().max().item()
        err = (got[:, sl] - want[:, sl]).abs().max().item()
        assert err <= scale * 2.0 ** -6, f"{name}: err {err:.4g} scale {scale:.4g}"
        if ref2 is not None:
            assert (got[:, sl] - ref2[:, sl]).abs().max().item() <= scale * 2.0 ** -6


@pytest.mark.parametrize("b,p,heads", [(3, 196, 32), (2, 49, 16), (1, 7, 8), (2, 1024, 8)])
def test_attnpool_query0_matches_torch(b, p, heads):
    """CLIP AttentionPool2d reduced to its token-0 output (M:89-127 via T:413): one query per image and head against the P patch
    keys plus the mean token, whose key / value are the means of the patch keys / values.  fp32 torch on the same bf16 K / V."""
    from lecb200 import ops
    c = heads * 64
    g = torch.Generator(device="cpu").manual_seed(p * 31 + heads)
    q = torch.randn((b, c), generator=g).cuda()
    kmat = torch.randn((b * p, c), generator=g).cuda().bfloat16()
    vmat = torch.randn((b * p, c), generator=g).cuda().bfloat16()
    out = ops.attnpool_query0(q, kmat, vmat, b, p, heads)
    torch.cuda.synchronize()
    kf = kmat.float().view(b, p, heads, 64)
    vf = vmat.float().view(b, p, heads, 64)
    keys = torch.cat([kf.mean(1, keepdim=True), kf], 1)                       # mean token first (M:96)
    vals = torch.cat([vf.mean(1, keepdim=True), vf], 1)
    s = torch.einsum("bhd,bthd->bht", q.view(b, heads, 64), keys) / 8.0
    ref = torch.einsum("bht,bthd->bhd", s.softmax(-1), vals).reshape(b, c)
    err = (out.float() - ref).abs().max().item()
    assert err < 2.0 ** -7 * max(1.0, ref.abs().max().item()), err          # bf16 output
