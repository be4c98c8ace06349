"""Kernel-level parity of lecb_head_aggregate (T:458-470 test / T:496-514 train) against oracle.restatement.aggregate in
float64 on the same dot products: every compile-time variant (evidence, maps, mask), class counts on both sides of the
lane-slot boundaries, and the inputs that force the kernel off its fast path (fixed softmax reference 0) onto the online
running-maximum sweep: un-normalised features and a spatial scale far above the reference's 50."""
import pytest
import torch

from oracle import restatement as R

pytestmark = pytest.mark.gpu


def _ops():
    from lecb200 import ops
    return ops


def _reference(dots, b, p, k, n_txt, ssq, mask, logit_scale, spatial_scale):
    d = dots.double().reshape(b, p, -1)
    if ssq is not None:
        d = d / ssq.double().reshape(b, p, 1).sqrt()
    pos = d[..., :k].permute(1, 0, 2)                              # [P,B,K]
    neg = d[..., k:2 * k].permute(1, 0, 2)
    evi = d[..., 2 * k:3 * k].permute(1, 0, 2) if n_txt == 3 else None
    if mask is not None:
        add = mask.reshape(b, p).t().double()[:, :, None] * (-10000.0)     # T:491-498
        neg = neg + add
        if evi is not None:
            evi = evi + add
    logits, neg_out = R.aggregate(neg, evi, spatial_scale, logit_scale)
    return logits, neg_out, pos


@pytest.mark.parametrize("k", [20, 64, 80, 128])
@pytest.mark.parametrize("n_txt", [2, 3])
@pytest.mark.parametrize("maps", [False, True])
def test_head_aggregate_matches_float64_oracle(k, n_txt, maps):
    ops = _ops()
    torch.manual_seed(k * 7 + n_txt)
    b, p, d = 5, 197, 64
    feats = torch.randn(b * p, d, device="cuda")
    txt = torch.nn.functional.normalize(torch.randn(n_txt * k, d, device="cuda"), dim=-1)
    dots = (feats @ txt.t()).contiguous()                                   # raw dots of UN-normalised rows
    ssq = (feats * feats).sum(-1).contiguous()
    out, neg, pos = ops.head_aggregate(dots, b, p, k, n_txt, row_sumsq=ssq, want_maps=maps)
    ref, ref_neg, ref_pos = _reference(dots, b, p, k, n_txt, ssq, None, 4.0, 50.0)
    assert float((out.double() - ref).abs().max()) < 2e-5                   # logits up to 4: fp32 accumulation level
    if maps:
        assert float((neg.double() - ref_neg).abs().max()) < 2e-6
        assert float((pos.double() - ref_pos).abs().max()) < 2e-6


@pytest.mark.parametrize("n_txt", [2, 3])
def test_head_aggregate_padding_mask(n_txt):
    """Caption path: rows behind the EOT carry the -10000 mask (weight exactly 0); whole warps see no live row."""
    ops = _ops()
    torch.manual_seed(3)
    b, p, k, d = 9, 77, 80, 64
    feats = torch.nn.functional.normalize(torch.randn(b * p, d, device="cuda"), dim=-1)
    txt = torch.nn.functional.normalize(torch.randn(n_txt * k, d, device="cuda"), dim=-1)
    dots = (feats @ txt.t()).contiguous()
    lens = torch.tensor([1, 2, 5, 8, 9, 20, 33, 76, 77], device="cuda")
    mask = (torch.arange(p, device="cuda")[None, :] >= lens[:, None]).to(torch.uint8).contiguous()
    out, neg, _ = ops.head_aggregate(dots, b, p, k, n_txt, row_mask=mask.reshape(-1), want_maps=True)
    ref, ref_neg, _ = _reference(dots, b, p, k, n_txt, None, mask, 4.0, 50.0)
    assert float((out.double() - ref).abs().max()) < 2e-5
    live = (mask.reshape(b, p).t() == 0)[:, :, None].expand_as(neg)
    assert float((neg.double() - ref_neg)[live].abs().max()) < 2e-6
    assert float(neg[~live].abs().max()) == 0.0                             # masked rows: not written (zero-filled by the op)


@pytest.mark.parametrize("case", ["unnormalised", "huge_scale", "tiny_scores"])
@pytest.mark.parametrize("n_txt", [2, 3])
def test_head_aggregate_leaves_the_fast_path_when_it_must(case, n_txt):
    """exp2(spatial score) with reference 0 would overflow / underflow here: the kernel must detect it from its result and
    redo the sweep with the running maximum, giving the oracle's numbers all the same."""
    ops = _ops()
    torch.manual_seed(11)
    b, p, k, d = 4, 196, 80, 32
    feats = torch.randn(b * p, d, device="cuda")
    txt = torch.nn.functional.normalize(torch.randn(n_txt * k, d, device="cuda"), dim=-1)
    scale = 50.0
    if case == "unnormalised":
        dots = (feats @ txt.t()).contiguous() * 1.5                         # |score| up to ~6: exp2(50*6*1.44) overflows
    elif case == "huge_scale":
        dots = (torch.nn.functional.normalize(feats, dim=-1) @ txt.t()).contiguous()
        scale = 400.0                                                       # 2^(+-577)
    else:
        dots = (torch.nn.functional.normalize(feats, dim=-1) @ txt.t()).contiguous() - 4.0    # every score ~ -4: 2^-288
    out, neg, _ = ops.head_aggregate(dots, b, p, k, n_txt, spatial_scale=scale, want_maps=True)
    ref, ref_neg, _ = _reference(dots, b, p, k, n_txt, None, None, 4.0, scale)
    assert torch.isfinite(out).all()
    # fp32 conditioning, not the kernel: the winner-take-all exponent gain*(neg - max) is a difference of numbers of size
    # gain*max (hundreds to thousands here), so one fp32 ulp of it is already 1e-4..1e-3 of the softmax weight
    tol = 2e-3 * max(1.0, float(ref.abs().max()))
    assert float((out.double() - ref).abs().max()) < tol
    assert float((neg.double() - ref_neg).abs().max()) < 2e-3 * max(1.0, float(ref_neg.abs().max()))


def test_head_aggregate_scale_property():
    """Headline-sized input (256 images x 196 patches): the result of an image does not depend on its batch."""
    ops = _ops()
    torch.manual_seed(5)
    b, p, k = 256, 196, 80
    dots = (torch.rand(b * p, 3 * k, device="cuda") * 2 - 1).contiguous()
    ssq = torch.ones(b * p, device="cuda")
    full, _, _ = ops.head_aggregate(dots, b, p, k, 3, row_sumsq=ssq, want_maps=False)
    part, _, _ = ops.head_aggregate(dots[17 * p:19 * p].contiguous(), 2, p, k, 3, row_sumsq=ssq[:2 * p], want_maps=False)
    assert torch.equal(full[17:19], part)
