"""Prompt-tuning step on the B200 vs reference-generated goldens: forward logits, both losses, and the
gradients of ctx / ctx_double / ctx_evidence (DenseCLIP.forward(None, captions) + loss.backward(),
T:473-545, T:805-815).  Tolerances: logits 1e-2 abs (north_star); loss 1 % relative; gradients 5 % of the
reference gradient's max-abs (bf16 operands through 12 transformer layers forward and backward)."""
import numpy as np
import pytest
import torch

from oracle import restatement as R

from . import _cases as C
from ._gpu_common import LOGIT_TOL, build_model

pytestmark = pytest.mark.gpu


def _fused_losses():
    from lecb200 import losses
    return losses


def test_losses_match_reference():
    L = _fused_losses()
    g = C.load("losses.npz")
    x0, y, yp = (torch.from_numpy(g[k]).cuda() for k in ("x", "y", "y_partial"))
    for name, fn in (("ranking_s1", lambda a: L.ranking_loss(a, y, scale_=1.0, margin_=1)),
                     ("ranking_s2", lambda a: L.ranking_loss(a, y)),
                     ("asl", lambda a: L.ASL_loss(a, y)),
                     ("dualcoop", lambda a: L.dualcoop_loss(a, None, yp))):
        a = x0.clone().requires_grad_(True)
        loss = fn(a)
        loss.backward()
        ref = float(g["loss_" + name])
        assert abs(loss.item() - ref) < 1e-4 * max(1.0, abs(ref)), name
        np.testing.assert_allclose(a.grad.cpu().numpy(), g["grad_" + name], atol=2e-6, rtol=1e-4, err_msg=name)
        np.testing.assert_array_equal(a.detach().cpu().numpy(), g["x"])       # no in-place scaling (U:86 quirk)


def _asl_float64(x, y, gamma_neg, gamma_pos, clip, eps, tp, tn, partial):
    """U:126-173 in float64 (focal weight outside the graph, U:162-170)."""
    x = x.double().requires_grad_(True)
    s = torch.sigmoid(x)
    pos, neg = (y > tp).double(), (y < tn).double()
    sneg = 1 - s
    if clip > 0:
        sneg = (sneg + clip).clamp(max=1)
    loss = pos * torch.log(s.clamp(min=eps)) + neg * torch.log(sneg.clamp(min=eps))
    with torch.no_grad():
        pt = s * pos + sneg * neg
        w = torch.pow(1 - pt, gamma_pos * pos + gamma_neg * neg)
    loss = -(loss * w).sum() / (x.shape[0] if partial else x.numel())
    loss.backward()
    return loss.detach(), x.grad


@pytest.mark.parametrize("tp,tn,gn", [(0.9, 0.9, 2.0), (0.9, -0.9, 2.0), (0.9, 0.9, 4.0), (0.3, 0.6, 2.0)])
def test_asl_kernel_variants_match_float64(tp, tn, gn):
    """The lean kernel (gamma 1 / 2, exclusive thresholds: flush-to-zero MUFU forms, one logarithm), the general fast-gamma kernel
    (overlapping thresholds: a target both positive and negative) and the powf kernel against the float64 formula, with
    saturated logits (sigmoid = 0 / 1 in fp32) and partial labels {-1, 0, 1} in the batch."""
    from lecb200 import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn((4096, 80), generator=g) * 3
    x[0, :8] = torch.tensor([-120.0, -90.0, -30.0, -17.0, 17.0, 30.0, 90.0, 120.0])
    y = torch.randint(-1, 2, (4096, 80), generator=g).float()
    y[1, :4] = torch.tensor([0.5, 0.45, 0.95, -0.95])
    x, y = x.cuda(), y.cuda()
    for partial in (False, True):
        loss, grad = ops.asl_fwd_bwd(x, y, gamma_neg=gn, gamma_pos=1.0, clip=0.05, eps=1e-8, thresh_pos=tp, thresh_neg=tn,
                                     partial=partial)
        want_l, want_g = _asl_float64(x, y, gn, 1.0, 0.05, 1e-8, tp, tn, partial)
        assert abs(loss.item() - want_l.item()) <= 2e-5 * max(1.0, abs(want_l.item())), (loss.item(), want_l.item())
        scale = want_g.abs().max().item()
        assert (grad.double() - want_g).abs().max().item() <= 2e-5 * scale
        assert torch.isfinite(grad).all()


def test_loss_kernels_scale():
    """Size-independent properties at a roofline-sized input: ASL of [2^18, 80] equals the mean of per-chunk
    losses; gradient rows only depend on their own row."""
    from lecb200 import ops
    torch.manual_seed(0)
    x = torch.randn((1 << 18, 80), device="cuda") * 2
    y = (torch.rand_like(x) < 0.04).float()
    loss, grad = ops.asl_fwd_bwd(x, y)
    parts = [ops.asl_fwd_bwd(x[i::4].contiguous(), y[i::4].contiguous())[0] for i in range(4)]
    assert abs(loss.item() - torch.stack(parts).mean().item()) < 1e-5
    l2, g2 = ops.asl_fwd_bwd(x[:1024].contiguous(), y[:1024].contiguous())
    torch.testing.assert_close(grad[:1024] * (x.shape[0] / 1024), g2, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("ev", [False, True])
@pytest.mark.parametrize("loss_name", ["ranking", "asl"])
def test_train_step_matches_reference(ev, loss_name):
    L = _fused_losses()
    c = C.train_case("rn50")
    g = c["gold"]
    head = dict(arch=c["arch"], sd=c["sd"], pl_state=c["pl_state"])
    model = build_model(head, use_evidence=ev)
    caps, y = c["captions"].cuda(), c["labels"].cuda()
    out = model(None, caps)
    assert len(out) == 6 and out[4] is None and out[5] is None
    logits, logits_local = out[0], out[1]
    if loss_name == "ranking":
        loss = L.ranking_loss(logits, y, scale_=1.0, margin_=1) + L.ranking_loss(logits_local, y, scale_=1.0, margin_=1)
    else:
        loss = L.ASL_loss(logits, y) + L.ASL_loss(logits_local, y)
    loss.backward()
    torch.cuda.synchronize()
    sfx = ("_ev" if ev else "") + "_" + loss_name
    s2 = "_ev" if ev else ""
    e1 = np.abs(logits.detach().cpu().numpy() - g["logits" + s2]).max()
    e2 = np.abs(logits_local.detach().cpu().numpy() - g["logits_local" + s2]).max()
    ref_loss = float(g["loss" + sfx])
    print(f"[train{sfx}] logits err {e1:.5f} local err {e2:.5f} loss {loss.item():.5f} vs {ref_loss:.5f}")
    assert e1 <= LOGIT_TOL and e2 <= LOGIT_TOL
    assert abs(loss.item() - ref_loss) <= 1e-2 * max(1.0, abs(ref_loss))
    np.testing.assert_allclose(out[3].detach().cpu().numpy(), g["text_features" + s2], atol=5e-3)
    # the caption tower runs up to the batch's last EOT only (exact under the causal mask); later positions are padding
    # with zero weight (T:491-498) and come back as zeros instead of the reference's pad-token features
    l_run = int(c["captions"].argmax(-1).max()) + 1
    feats = out[2].detach().cpu().numpy()
    assert feats.shape[0] == c["captions"].shape[1]
    np.testing.assert_allclose(feats[:l_run, :2], g["seq_feats" + s2][:l_run], atol=5e-3)
    assert l_run == feats.shape[0] or float(np.abs(feats[l_run:]).max()) == 0.0
    pl = model.prompt_learner
    for pname in ("ctx", "ctx_double", "ctx_evidence"):
        gref = g[f"grad_{pname}" + sfx]
        got = getattr(pl, pname).grad
        if bool(g[f"gradnone_{pname}" + sfx]):
            assert got is None or float(got.abs().max()) == 0.0, pname
            continue
        scale = np.abs(gref).max()
        err = np.abs(got.cpu().numpy() - gref).max() / scale
        cos = float((got.cpu().flatten() @ torch.from_numpy(gref).flatten()) / (got.cpu().norm() * np.linalg.norm(gref)))
        print(f"[train{sfx}] grad {pname}: max err / max ref = {err:.4f}, cosine = {cos:.5f}")
        # ASL is smooth: 5 % of the largest reference entry.  The ranking loss is piecewise linear (U:85-93): a logit that
        # differs from the reference's by the bf16-level 7e-3 flips every hinge within that distance of the margin, and each
        # flip moves dlogits by a whole unit; on top of that two runs of the backward itself differ by up to 1.8 % of max
        # (fp32 atomics feeding bf16 roundings).  Measured over repeated runs: 2.7-5.0 % (ranking), 1-2.5 % (ASL);
        # the direction (cosine) is the stable criterion and keeps the same gate for both
        gate = 8e-2 if loss_name == "ranking" else 5e-2
        assert err < gate and cos > 0.999, (pname, err, cos)


def test_caption_bank_builder_matches_oracle(tmp_path):
    """bank rows == unit EOT features of the oracle text tower; pickle round trip; retrieval consumes it."""
    from lecb200 import bank as B
    c = C.train_case("rn50")
    head = dict(arch=c["arch"], sd=c["sd"], pl_state=c["pl_state"])
    model = build_model(head, use_evidence=False)
    caps = C.synth.captions(24, 5, vocab=c["arch"].vocab_size)
    bank = B.build_caption_bank(model.text_encoder, caps, batch_size=16)
    assert bank.dtype == torch.float16 and tuple(bank.shape) == (24, c["arch"].embed_dim)
    with torch.no_grad():
        ref = R.text_encode(c["sd"], R.embed_tokens(c["sd"], caps), caps.argmax(-1), c["arch"].transformer_heads)
        ref = ref / ref.norm(dim=-1, keepdim=True)
    err = (bank.float().cpu() - ref).abs().max().item()
    assert err < 5e-3, err                          # unit vectors of dim 1024 through a bf16 tower
    path = str(tmp_path / "bank.pkl")
    B.save_caption_bank(path, bank)
    import pickle
    with open(path, "rb") as f:
        raw = pickle.load(f)                        # what Caption_distill_double.py:35-36 does
    assert isinstance(raw, torch.Tensor) and raw.device.type == "cpu" and torch.equal(raw, bank.cpu())
    assert torch.equal(B.load_caption_bank(path), bank)
