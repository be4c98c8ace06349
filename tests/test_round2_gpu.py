"""Round-2 GPU parity: the config switches round 1 left unpinned (TRAIN.ema, CSC, IF_LEARN_SCALE, co-occurrence ranking),
the rewritten / new kernels (ranking, KL, multi-tensor updates, fused retrieval top-10, uint8 stem, window crop + resize),
and the drop-in wiring (DDP wrapper, 5-tuple without a caption bank).  Everything goes through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import synth

from . import _cases as C
from ._gpu_common import LOGIT_TOL, build_model
from .test_oracle_ext import ext_case, pl_state

pytestmark = pytest.mark.gpu


def _losses():
    from lecb200 import losses
    return losses


def _train_model(c, ev, csc=False, **kw):
    head = dict(arch=c["arch"], sd=c["sd"], pl_state=pl_state(c, csc=csc))
    return build_model(head, use_evidence=ev, csc=csc, **kw)


def _check_grads(model, g, sfx, names=("ctx", "ctx_double", "ctx_evidence"), tol=5e-2, min_cos=0.999):
    from oracle.make_golden import CSC_ROWS
    for pname in names:
        gref = g[f"grad_{pname}" + sfx]
        got = getattr(model.prompt_learner, pname).grad
        if bool(g[f"gradnone_{pname}" + sfx]):
            assert got is None or float(got.abs().max()) == 0.0, pname
            continue
        got = got.detach().float().cpu().numpy()
        if got.ndim == 3 and got.shape[0] != gref.shape[0]:
            norms = np.linalg.norm(got.reshape(got.shape[0], -1), axis=1)
            np.testing.assert_allclose(norms, g[f"gradnorm_{pname}" + sfx], rtol=5e-2, atol=1e-3 * g[f"gradnorm_{pname}" + sfx].max())
            got = got[list(CSC_ROWS)]
        scale = np.abs(gref).max()
        err = np.abs(got - gref).max() / scale
        cos = float((got.flatten() @ gref.flatten()) / (np.linalg.norm(got) * np.linalg.norm(gref)))
        print(f"[{sfx}] grad {pname}: max err / max ref = {err:.4f}, cosine = {cos:.5f}")
        assert err < tol and cos > min_cos, (pname, err, cos)


# ---------------------------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------------------------
def test_losses_ext_match_reference():
    L = _losses()
    g = C.load("losses_ext.npz")
    x0, xm, y, p = (torch.from_numpy(g[k]).cuda() for k in ("x", "xm", "y", "cooc_p"))
    for name, fn in (("cooc_s1", lambda a: L.ranking_loss_with_cooccurrence(a, y, p, scale_=1.0, margin_=1)),
                     ("cooc_s2", lambda a: L.ranking_loss_with_cooccurrence(a, y, p)),
                     ("kl", lambda a: L.kl_softmax(a, xm)),
                     ("kl_x10000", lambda a: L.kl_softmax(a, xm, 10000.0))):
        a = x0.clone().requires_grad_(True)
        loss = fn(a)
        loss.backward()
        ref = float(g["loss_" + name])
        assert abs(loss.item() - ref) < 1e-4 * max(1.0, abs(ref)), (name, loss.item(), ref)
        gmax = np.abs(g["grad_" + name]).max()
        np.testing.assert_allclose(a.grad.cpu().numpy(), g["grad_" + name], atol=2e-6 * max(1.0, gmax), rtol=1e-4, err_msg=name)
        np.testing.assert_array_equal(a.detach().cpu().numpy(), g["x"])


def test_resample_loss_matches_reference():
    """`ResampleLoss` (trainers/dbl.py:263-445, LOSSFUNC 'dbl' at T:818-841) through the fused kernel, same constructor keywords."""
    L = _losses()
    from oracle.make_golden import RESAMPLE_CONFIGS
    g = C.load("resample_loss.npz")
    x0, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    for name, kw in RESAMPLE_CONFIGS.items():
        fn = L.ResampleLoss(class_freq=g["class_freq"], neg_class_freq=g["neg_class_freq"], **kw)
        a = x0.clone().requires_grad_(True)
        loss = fn(a, y)
        loss.backward()
        ref = float(g["loss_" + name])
        assert abs(loss.item() - ref) < 1e-5 * max(1.0, abs(ref)), (name, loss.item(), ref)
        np.testing.assert_allclose(a.grad.cpu().numpy(), g["grad_" + name], atol=2e-7, rtol=2e-4, err_msg=name)
        np.testing.assert_array_equal(a.detach().cpu().numpy(), g["x"])
    # size-independent property at a roofline-sized input: the mean over a batch is the mean of the means of its equal parts
    from lecb200 import ops
    torch.manual_seed(4)
    xb = torch.randn((1 << 16, 80), device="cuda") * 2
    yb = (torch.rand_like(xb) < 0.05).float()
    fi = (1.0 / torch.from_numpy(g["class_freq"])).cuda().contiguous()
    whole, gw = ops.resample_bce_fwd_bwd(xb, yb, fi, None, 0.1, 10.0, 0.2)
    parts = [ops.resample_bce_fwd_bwd(xb[i::4].contiguous(), yb[i::4].contiguous(), fi, None, 0.1, 10.0, 0.2)[0] for i in range(4)]
    assert abs(whole.item() - torch.stack(parts).mean().item()) < 1e-5 * abs(whole.item())
    _, g2 = ops.resample_bce_fwd_bwd(xb[:1024].contiguous(), yb[:1024].contiguous(), fi, None, 0.1, 10.0, 0.2)
    torch.testing.assert_close(gw[:1024] * (xb.shape[0] / 1024), g2, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("k", [80, 33, 200, 6])
@pytest.mark.parametrize("soft", [False, True])
def test_ranking_kernel_matches_dense_formula(k, soft):
    """The sparse-list kernel against the reference's dense [B,K,K] expression (U:85-93), incl. non-binary targets, rows
    without positives and rows with every label set."""
    from lecb200 import ops
    torch.manual_seed(k)
    b = 37
    x = torch.randn((b, k), device="cuda") * 2
    y = (torch.rand((b, k), device="cuda") < 0.1).float()
    y[0] = 0.0
    y[1] = 1.0
    if soft:
        y = y * torch.rand_like(y) + 0.05 * (torch.rand_like(y) < 0.1).float()
    w = torch.rand((k, k), device="cuda") + 0.5
    for cooc in (False, True):
        a = x.clone().requires_grad_(True)
        s = a * 2.0
        d = (1.0 - s[:, None, :] + s[:, :, None]).clamp_min(0)
        if cooc:
            d = d * w
        ref = (d * y[:, None, :] * (1 - y[:, :, None])).sum((-1, -2)).mean()
        ref.backward()
        loss, grad = (ops.ranking_cooc_fwd_bwd(x, y, w, 2.0, 1.0) if cooc else ops.ranking_fwd_bwd(x, y, 2.0, 1.0))
        assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item())), (cooc, loss.item(), ref.item())
        torch.testing.assert_close(grad, a.grad, rtol=1e-4, atol=1e-5 * max(1.0, float(a.grad.abs().max())))


def test_ranking_kernel_scale_property():
    """Roofline-sized input [2^18, 80]: loss = mean of per-chunk losses, gradient rows depend on their own row only."""
    from lecb200 import ops
    torch.manual_seed(1)
    x = torch.randn((1 << 18, 80), device="cuda") * 2
    y = (torch.rand_like(x) < 0.04).float()
    loss, grad = ops.ranking_fwd_bwd(x, y, 1.0, 1.0)
    parts = [ops.ranking_fwd_bwd(x[i::4].contiguous(), y[i::4].contiguous(), 1.0, 1.0)[0] for i in range(4)]
    assert abs(loss.item() - torch.stack(parts).mean().item()) < 1e-4 * abs(loss.item())
    _, g2 = ops.ranking_fwd_bwd(x[:1024].contiguous(), y[:1024].contiguous(), 1.0, 1.0)
    torch.testing.assert_close(grad[:1024] * (x.shape[0] / 1024), g2, rtol=1e-5, atol=1e-7)


def test_kl_kernel_matches_torch():
    from lecb200 import ops
    torch.manual_seed(2)
    for b, k in ((64, 80), (7, 6), (300, 200)):
        x = torch.randn((b, k), device="cuda") * 3
        xm = x + torch.randn_like(x) * 0.5
        a = x.clone().requires_grad_(True)
        ref = torch.nn.KLDivLoss(reduction="batchmean")(torch.log_softmax(a, -1), torch.softmax(xm, -1)) * 7.0
        ref.backward()
        loss, grad = ops.kl_softmax_fwd_bwd(x, xm, 7.0)
        assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
        torch.testing.assert_close(grad, a.grad, rtol=1e-4, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# multi-tensor updates
# ---------------------------------------------------------------------------------------------------------------------
def test_multi_tensor_kernels_match_torch():
    from lecb200 import ops
    torch.manual_seed(3)
    shapes = [(16, 512), (80, 16, 512), (16, 512), (), (), ()]
    live = [torch.randn(s, device="cuda") for s in shapes]
    twin = [torch.randn(s, device="cuda") for s in shapes]
    want = [t * 0.995 + l * (1.0 - 0.995) for l, t in zip(live, twin)]
    ops.ema_update(live, twin, 0.995)
    for w, t in zip(want, twin):
        assert torch.equal(t, w)                      # same rounding points as ATen
    grads = [torch.randn(s, device="cuda") for s in shapes]
    grads[2] = None                                       # ctx_evidence without gradient (use_evidence off)
    flat = ops.pack_f32(grads, live)
    ref = torch.cat([(torch.zeros_like(l) if g is None else g).reshape(-1) for g, l in zip(grads, live)])
    assert torch.equal(flat, ref)
    outs = [torch.empty(s, device="cuda") for s in shapes]
    ops.unpack_scale_f32(flat, outs, 0.125)
    off = 0
    for o in outs:
        assert torch.equal(o.reshape(-1), ref[off:off + o.numel()] * 0.125)
        off += o.numel()
    # SGD with momentum + weight decay over three steps == torch.optim.SGD
    params = [torch.randn(s, device="cuda") for s in shapes]
    tparams = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.SGD(tparams, lr=0.002, momentum=0.9, weight_decay=5e-4)
    bufs = [torch.zeros_like(p) for p in params]
    for step in range(3):
        gs = [torch.randn(s, device="cuda") for s in shapes]
        for tp, g_ in zip(tparams, gs):
            tp.grad = g_.clone() * 0.5
        opt.step()
        ops.sgd_step(ops.pack_f32(gs, params), params, bufs, 0.002, 0.9, 5e-4, grad_scale=0.5)
        for tp, p_ in zip(tparams, params):
            torch.testing.assert_close(p_, tp.detach(), rtol=1e-6, atol=1e-7)


# ---------------------------------------------------------------------------------------------------------------------
# prompt-tuning switches vs reference goldens
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ev", [False, True])
def test_ema_step_matches_reference(ev):
    """TRAIN.ema=True (a14): 6-tuple with the twin's logits after the momentum update, ranking + EMA-KL loss, gradients."""
    L = _losses()
    c = ext_case("rn50")
    g = c["gold"]
    sfx = "_ema" + ("_ev" if ev else "")
    model = _train_model(c, ev, ema=True)
    tw = pl_state(c, twin=True)
    with torch.no_grad():
        for pname in ("ctx", "ctx_double", "ctx_evidence"):
            getattr(model.prompt_learner_m, pname).copy_(tw[pname])
    caps, y = c["captions"].cuda(), c["labels"].cuda()
    out = model(None, caps)
    assert len(out) == 6 and out[4] is not None and out[5] is not None
    for pname in ("ctx", "ctx_double", "ctx_evidence"):           # _momentum_update ran before the twin's forward (T:518)
        np.testing.assert_allclose(getattr(model.prompt_learner_m, pname).detach().cpu().numpy(), g[f"twin_{pname}" + sfx], atol=1e-7)
        assert not getattr(model.prompt_learner_m, pname).requires_grad
    errs = [np.abs(out[i].detach().cpu().numpy() - g[k + sfx]).max() for i, k in
            ((0, "logits"), (1, "logits_local"), (4, "logits_m"), (5, "logits_local_m"))]
    print(f"[{sfx}] logit errors {errs}")
    assert max(errs) <= LOGIT_TOL
    r_loss = L.ranking_loss(out[0], y, scale_=1.0, margin_=1) + L.ranking_loss(out[1], y, scale_=1.0, margin_=1)
    ema_loss = L.ema_consistency_loss(out[0], out[4], out[1], out[5])
    assert abs(r_loss.item() - float(g["r_loss" + sfx])) <= 1e-2 * max(1.0, abs(float(g["r_loss" + sfx])))
    # kernel vs the oracle's expression at the product's own logits (tight), and vs the reference's number (the 10000 x
    # local KL sees the bf16-level logit error of live and twin: 5 %)
    # (in float64: the 10000 x local term is a sum of q * (log q - log p) with log differences of ~1e-3 between logs of ~4,
    # so ANY fp32 evaluation — the kernel's, torch's — carries ~1e-3 of relative rounding noise)
    with torch.no_grad():
        want = R.ema_loss(*(out[i].detach().double().cpu() for i in (0, 4, 1, 5)))
    print(f"[{sfx}] ema_loss kernel {ema_loss.item():.6f} vs float64 oracle at the same logits {want.item():.6f}")
    assert abs(ema_loss.item() - want.item()) <= 3e-3 * max(1.0, abs(want.item()))
    ref_ema = float(g["ema_loss" + sfx])
    print(f"[{sfx}] ema_loss {ema_loss.item():.5f} vs reference {ref_ema:.5f}")
    assert abs(ema_loss.item() - ref_ema) <= 5e-2 * max(1.0, abs(ref_ema))
    (r_loss + ema_loss).backward()
    torch.cuda.synchronize()
    # the 10000 x local KL multiplies the bf16-level error of (p_live - p_twin): a little looser than the plain step
    _check_grads(model, g, sfx, tol=8e-2, min_cos=0.998)
    for p in model.prompt_learner_m.parameters():
        assert p.grad is None


@pytest.mark.parametrize("ev", [False, True])
def test_learnable_scale_matches_reference(ev):
    """TRAIN.IF_LEARN_SCALE=True: logit scale exp(temperature) = e^3, and d loss / d temperature (hand-written d_scale)."""
    L = _losses()
    c = ext_case("rn50")
    g = c["gold"]
    sfx = "_scale" + ("_ev" if ev else "")
    model = _train_model(c, ev, learn_scale=True)
    caps, y = c["captions"].cuda(), c["labels"].cuda()
    out = model(None, caps)
    out[0].retain_grad()
    out[1].retain_grad()
    scale = float(np.exp(3.0)) / 4.0                # logits are 5x larger than with the fixed scale 4: tolerance scales along
    e1 = np.abs(out[0].detach().cpu().numpy() - g["logits" + sfx]).max()
    e2 = np.abs(out[1].detach().cpu().numpy() - g["logits_local" + sfx]).max()
    print(f"[{sfx}] logits err {e1:.5f} local err {e2:.5f}")
    assert e1 <= LOGIT_TOL * scale and e2 <= LOGIT_TOL * scale
    loss = L.ranking_loss(out[0], y, scale_=1.0, margin_=1) + L.ranking_loss(out[1], y, scale_=1.0, margin_=1)
    loss.backward()
    torch.cuda.synchronize()
    ref = float(g["loss" + sfx])
    assert abs(loss.item() - ref) <= 1e-2 * max(1.0, abs(ref))
    gt, got = float(g["grad_temperature" + sfx]), float(model.prompt_learner.temperature.grad)
    # logits = exp(temperature) * (...)  =>  d loss / d temperature = sum(dlogits * logits) + sum(dlocal * logits_local): a small
    # residual of large cancelling terms.  (1) the hand-written d_scale reproduces that expression on the product's own
    # tensors; (2) against the reference the error is bounded relative to the sum of the absolute terms (what a 1e-2-accurate
    # logit can move it by)
    terms = torch.cat([(out[0].grad * out[0].detach()).flatten(), (out[1].grad * out[1].detach()).flatten()]).double()
    mass = float(terms.abs().sum())
    print(f"[{sfx}] temperature grad {got:.5f} vs reference {gt:.5f}; expression on own tensors {float(terms.sum()):.5f}; sum |terms| {mass:.2f}")
    assert abs(got - float(terms.sum())) <= 1e-4 * mass
    assert abs(got - gt) <= 1e-2 * mass
    _check_grads(model, g, sfx)


def test_csc_step_matches_reference():
    """TRAINER.Caption.CSC=True: class-specific ctx / ctx_double [80,16,512] (the 5.3 MB all-reduce case)."""
    L = _losses()
    c = ext_case("rn50")
    g = c["gold"]
    model = _train_model(c, True, csc=True)
    assert tuple(model.prompt_learner.ctx.shape) == (80, 16, 512) and tuple(model.prompt_learner.ctx_evidence.shape) == (16, 512)
    caps, y = c["captions"].cuda(), c["labels"].cuda()
    out = model(None, caps)
    np.testing.assert_allclose(out[3].detach().cpu().numpy(), g["text_features_csc_ev"], atol=5e-3)
    e1 = np.abs(out[0].detach().cpu().numpy() - g["logits_csc_ev"]).max()
    e2 = np.abs(out[1].detach().cpu().numpy() - g["logits_local_csc_ev"]).max()
    rng = np.abs(g["logits_local_csc_ev"]).max()
    print(f"[_csc_ev] logits err {e1:.5f} local err {e2:.5f} (range {rng:.3f})")
    assert e1 <= LOGIT_TOL
    # With class-specific contexts the bf16-level errors of the 80 negative-prompt features are independent, and the winner-
    # take-all softmax (scale 50 x (max + 1), T:508) amplifies them ~75x instead of cancelling them as a common mode (the
    # generic-context cases stay inside 1e-2): 3e-2 absolute and 1.5 % of the output range here
    assert e2 <= 3 * LOGIT_TOL and e2 <= 1.5e-2 * rng
    loss = L.ranking_loss(out[0], y, scale_=1.0, margin_=1) + L.ranking_loss(out[1], y, scale_=1.0, margin_=1)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss_csc_ev"])) <= 1e-2 * max(1.0, abs(float(g["loss_csc_ev"])))
    _check_grads(model, g, "_csc_ev")


def test_cooccurrence_step_matches_reference():
    """LOSSFUNC 'ranking_with_cooccurrence' (T:842-850) with the reference's freq_stats prior."""
    L = _losses()
    c = ext_case("rn50")
    g = c["gold"]
    model = _train_model(c, False)
    caps, y = c["captions"].cuda(), c["labels"].cuda()
    p = torch.from_numpy(g["cooc_p"]).cuda()
    out = model(None, caps)
    loss = L.ranking_loss_with_cooccurrence(out[0], y, p, scale_=1.0, margin_=1) + \
        L.ranking_loss_with_cooccurrence(out[1], y, p, scale_=1.0, margin_=1)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss_cooc"])) <= 1e-2 * max(1.0, abs(float(g["loss_cooc"])))
    _check_grads(model, g, "_cooc")


# ---------------------------------------------------------------------------------------------------------------------
# retrieval
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,b", [(5000, 1024, 8), (2003, 512, 130), (70001, 1024, 256)])
def test_fused_retrieval_matches_unfused_and_oracle(n, d, b):
    """lecb_gemm_topk10 + merge == the explicit similarity matrix path == the oracle; N not a multiple of 8 included."""
    from lecb200 import retrieval
    bank = synth.caption_bank(n, d, 11).cuda()
    torch.manual_seed(n)
    g = torch.nn.functional.normalize(torch.randn((b, d), device="cuda"), dim=-1)
    g[: min(b, 4)] = torch.nn.functional.normalize(bank[[5, n - 1, n // 2, 17][: min(b, 4)]].float() + 0.01 * g[: min(b, 4)], dim=-1)
    g_add, vals = retrieval.retrieve_mean(g, bank)
    sim = g.double() @ bank.double().t()
    ref_vals, ref_idx = sim.topk(10, -1)
    err = (vals.double() - ref_vals).abs().max().item()
    print(f"retrieval n={n} d={d} b={b}: top-10 score max err {err:.3g}")
    assert err < 2e-5                                  # fp32 TMEM accumulation over K = 2 x D terms of a sum near 1 (planted rows)
    ref_add = bank[ref_idx.reshape(-1)].reshape(b, 10, d).float().mean(1).half().float()
    gaps = (ref_vals[:, 9] - sim.topk(11, -1)[0][:, 10])
    ok = gaps > 1e-6                                      # rows whose 10th / 11th scores are separated: the set is unique
    assert ok.float().mean() > 0.9
    assert (g_add[ok] - ref_add[ok]).abs().max().item() < 1e-3
    if n % 8 == 0:
        g2, v2, i2 = retrieval.retrieve_mean_unfused(g, bank, return_idx=True)
        torch.testing.assert_close(vals, v2, rtol=0, atol=5e-6)        # different k-block order of the two GEMM schedules
        assert (g_add[ok] - g2[ok]).abs().max().item() < 1e-3


# ---------------------------------------------------------------------------------------------------------------------
# uint8 stem + window crop / resize
# ---------------------------------------------------------------------------------------------------------------------
def test_stem_u8_equals_float_path():
    """lecb_stem_conv1_u8 on raw pixels == lecb_stem_conv1 on ToTensor + Normalize of the same pixels, bit for bit."""
    from lecb200 import ops
    torch.manual_seed(5)
    for (b, h, w) in ((3, 64, 96), (2, 224, 224), (1, 50, 38)):
        u8 = torch.randint(0, 256, (b, h, w, 3), dtype=torch.uint8)
        mean = torch.tensor(ops.CLIP_PIXEL_MEAN, dtype=torch.float32)
        std = torch.tensor(ops.CLIP_PIXEL_STD, dtype=torch.float32)
        x = ((u8.float() / 255.0 - mean) / std).permute(0, 3, 1, 2).contiguous()      # torchvision ToTensor + Normalize
        w27 = (torch.randn((27, 32)) * 0.2).cuda()
        bias = (torch.randn((32,)) * 0.1).cuda()
        a = ops.stem_conv1(x.cuda(), w27, bias)
        bq = ops.stem_conv1_u8(u8.cuda(), w27, bias)
        assert torch.equal(a, bq), (b, h, w, (a.float() - bq.float()).abs().max().item())
        # and both equal the fp32 convolution of the bf16-rounded operands
        wt = w27.view(3, 3, 3, 32).permute(3, 0, 1, 2).bfloat16().float()
        ref = torch.nn.functional.conv2d(x.cuda().bfloat16().float(), wt, bias, stride=2, padding=1).relu().permute(0, 2, 3, 1)
        assert (a.float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


def test_window_crop_resize_bit_exact():
    """Sliding windows of a random image: crop + resize on the GPU == numpy crop + oracle/pil_resize.py (itself bit-exact
    against Pillow), byte for byte; the float output == the reference's Resize -> ToTensor -> Normalize to 1e-6."""
    from lecb200 import ops
    from lecb200 import windows as WN
    from oracle import pil_resize as PR
    rng = np.random.default_rng(3)
    for (h, w, size, scales) in ((375, 500, 448, (2, 3)), (240, 180, 224, (2,))):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        img[: h // 4] = np.where(rng.random((h // 4, w, 1)) < 0.5, 0, 255)        # saturating edges: the clamp matters
        wins = [WN.whole_image(h, w)] + [x for s in scales for x in WN.sliding_windows(h, w, s)]
        u8, f32 = WN.crop_resize(torch.from_numpy(img).cuda(), wins, size, want_u8=True, want_f32=True)
        u8, f32 = u8.cpu().numpy(), f32.cpu().numpy()
        bad = 0
        for i, win in enumerate(wins):
            rows, cols = WN.source_rows_cols(win, h, w)
            crop = np.ascontiguousarray(img[rows][:, cols])
            want = PR.resize_u8(crop, size, size, "bicubic")
            if not np.array_equal(u8[i], want):
                bad += 1
                continue
            ref_f = PR.test_transform(crop, (size, size), ops.CLIP_PIXEL_MEAN, ops.CLIP_PIXEL_STD)
            assert np.abs(f32[i] - ref_f).max() <= 1e-6
        assert bad == 0, f"{bad} of {len(wins)} windows differ from Pillow's bytes"


def test_u8_batch_through_the_model_equals_float_batch():
    """DenseCLIPB200(image_u8) == DenseCLIPB200(normalised float image): same logits exactly (same bf16 stem operand)."""
    from lecb200 import ops
    c = C.head_case("small")
    model = build_model(c, use_evidence=True)
    res = c["arch"].image_resolution
    torch.manual_seed(9)
    u8 = torch.randint(0, 256, (4, res, res, 3), dtype=torch.uint8)
    mean = torch.tensor(ops.CLIP_PIXEL_MEAN, dtype=torch.float32)
    std = torch.tensor(ops.CLIP_PIXEL_STD, dtype=torch.float32)
    x = ((u8.float() / 255.0 - mean) / std).permute(0, 3, 1, 2).contiguous()
    a = model(x.cuda(), if_test=True)
    b = model(u8.cuda(), if_test=True)
    for t1, t2 in zip(a[:4], b[:4]):              # same bf16 stem operand; fp32 atomics (row sums of squares) reorder: 1e-5
        assert (t1 - t2).abs().max().item() <= 1e-5
    assert tuple(a[4].shape) == (4, 10) and float(a[4].abs().max()) == 0.0        # no caption bank: zeros, not None (T:645)


# ---------------------------------------------------------------------------------------------------------------------
# planted prototypes at the headline shape, evidence / WTA path (logits of order 1-4 instead of 0.1)
# ---------------------------------------------------------------------------------------------------------------------
def test_planted_prototypes_rn101_448_evidence():
    c = C.head_case("rn101_448")
    arch = c["arch"]
    model = build_model(c, use_evidence=True)
    img = c["image"][:2]
    with torch.no_grad():
        feat = R.rn_trunk(c["sd"], img, arch.vision_layers)
        local = R.local_features(c["sd"], feat)                              # [P,B,D]
        g = R.attnpool_global(c["sd"], feat, arch.vision_width * 32 // 64)
    k, d = 80, arch.embed_dim
    gen = torch.Generator().manual_seed(17)
    lu = local / local.norm(dim=-1, keepdim=True)
    gu = g / g.norm(dim=-1, keepdim=True)
    # prototypes: class j of every prompt set points at a real patch / global feature plus noise -> similarities up to ~0.9
    t_pos, t_neg, t_evi = (torch.randn((k, d), generator=gen) * 0.05 for _ in range(3))
    for j in range(k):
        pch = int(torch.randint(0, lu.shape[0], (1,), generator=gen))
        bi = j % lu.shape[1]
        t_neg[j] += lu[pch, bi] * (0.5 + 0.5 * (j % 3 == 0))
        t_evi[j] += lu[(pch * 7 + 3) % lu.shape[0], bi]
        t_pos[j] += gu[bi] * (1.0 if j % 2 == 0 else 0.3)
    t_pos, t_neg, t_evi = (t / t.norm(dim=-1, keepdim=True) for t in (t_pos, t_neg, t_evi))
    model.prompt_text_features = {"text_features": t_pos.cuda(), "text_features_neg": t_neg.cuda(),
                                  "text_features_evidence": t_evi.cuda()}
    out = model(img.cuda(), if_test=True)
    with torch.no_grad():
        ref = R.head_test(g, local, t_pos, t_neg, t_evi)
    assert ref[0].abs().max() > 2.0 and ref[1].abs().max() > 0.5, (ref[0].abs().max(), ref[1].abs().max())
    errs = [(o.float().cpu() - r).abs().max().item() for o, r in zip(out[:4], ref[:4])]
    print(f"planted rn101_448 evidence: ref absmax logits {ref[0].abs().max():.3f} local {ref[1].abs().max():.3f}; errors {errs}")
    assert max(errs) <= LOGIT_TOL, errs
    from ._gpu_common import topk_sets_match
    assert topk_sets_match(out[0].cpu().numpy(), ref[0].numpy(), 5, LOGIT_TOL)
    assert topk_sets_match(out[1].cpu().numpy(), ref[1].numpy(), 5, LOGIT_TOL)


# ---------------------------------------------------------------------------------------------------------------------
# drop-in wiring: the reference wraps the model in DistributedDataParallel(find_unused_parameters=True) (T:786-787)
# ---------------------------------------------------------------------------------------------------------------------
def test_ddp_wrapped_step_equals_unwrapped(tmp_path):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    L = _losses()
    c = C.train_case("rn50")
    head = dict(arch=c["arch"], sd=c["sd"], pl_state=c["pl_state"])
    caps, y = c["captions"].cuda(), c["labels"].cuda()

    def grads(wrap):
        model = build_model(head, use_evidence=False)
        for name, p in model.named_parameters():                    # T:763-765
            if "prompt_learner" not in name:
                p.requires_grad_(False)
        net = DDP(model, device_ids=[0], output_device=0, find_unused_parameters=True) if wrap else model
        opt = torch.optim.SGD(model.prompt_learner.parameters(), lr=0.002, momentum=0.9)
        out = net(None, caps)
        loss = L.ranking_loss(out[0], y, scale_=1.0, margin_=1) + L.ranking_loss(out[1], y, scale_=1.0, margin_=1)
        opt.zero_grad()
        loss.backward()
        opt.step()
        pl = model.prompt_learner
        return loss.item(), [None if p.grad is None else p.grad.clone() for p in (pl.ctx, pl.ctx_double, pl.ctx_evidence)], \
            [p.detach().clone() for p in (pl.ctx, pl.ctx_double)]

    own = not dist.is_initialized()
    if own:
        dist.init_process_group("nccl", init_method=f"file://{tmp_path}/rdzv", rank=0, world_size=1)
    try:
        l1, g1, p1 = grads(False)
        l0, g0, p0 = grads(False)
        l2, g2, p2 = grads(True)
    finally:
        if own:
            dist.destroy_process_group()
    assert abs(l1 - l2) < 1e-5 * max(1.0, abs(l1))

    def closeness(a, b):
        err = float((a - b).abs().max() / a.abs().max())
        cos = float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))
        return err, cos

    # The backward is not bit-reproducible (fp32 atomics in the small transposed GEMMs feed bf16 roundings), so two
    # identical un-wrapped runs set the yardstick and the wrapped run must agree with them as well as they agree with
    # each other — and well inside the tolerance the gradients are held to against the reference (5 % of max, cos 0.999)
    for a, a0, b in zip(g1, g0, g2):
        if a is None:
            assert b is None or float(b.abs().max()) == 0.0          # DDP materialises unused gradients as zeros
            continue
        e_rep, c_rep = closeness(a, a0)
        e_ddp, c_ddp = closeness(a, b)
        print(f"DDP-wrapped vs plain: err/max {e_ddp:.2e} cos {c_ddp:.6f}   (plain vs plain: {e_rep:.2e}, {c_rep:.6f})")
        assert e_ddp <= max(2e-2, 3 * e_rep) and c_ddp > 0.9995
    for a, b, ga in zip(p1, p2, g1):             # one SGD step: lr x the gradient difference allowed above
        assert float((a - b).abs().max()) <= 0.002 * 5e-2 * float(ga.abs().max()) + 1e-7


def test_window_pipeline_matches_host_pipeline():
    """pipeline.score_image_with_windows (crop + resize + score + fuse on the GPU) == the reference's flow re-enacted with the
    oracle's Pillow-exact resize on the host, the same model for the scoring and the oracle's aggregation rule (T:641-673)."""
    from lecb200 import pipeline
    from lecb200 import windows as WN
    from oracle import pil_resize as PR
    from lecb200 import ops
    c = C.head_case("small")
    model = build_model(c, use_evidence=True)
    size = c["arch"].image_resolution
    rng = np.random.default_rng(11)
    h, w = 150, 200
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    got = pipeline.score_image_with_windows(model, torch.from_numpy(img).cuda(), size, multi_scale=(2, 3), chunk=64)
    wins = [WN.whole_image(h, w)] + [x for s in (2, 3) for x in WN.sliding_windows(h, w, s)]
    assert got["n_windows"] == len(wins) - 1
    host = []
    for win in wins:
        rows, cols = WN.source_rows_cols(win, h, w)
        host.append(PR.test_transform(np.ascontiguousarray(img[rows][:, cols]), (size, size), ops.CLIP_PIXEL_MEAN, ops.CLIP_PIXEL_STD))
    x = torch.from_numpy(np.stack(host)).cuda()
    o, o_pos = [], []
    for i in range(0, x.shape[0], 64):
        r = model(x[i:i + 64], if_test=True)
        o.append(r[0])
        o_pos.append(r[1])
    o, o_pos = torch.cat(o).cpu(), torch.cat(o_pos).cpu()
    want = R.aggregate_blocks(o[:1], o[1:].unsqueeze(0), 0.3, 1.4)
    want_pos = R.aggregate_blocks(o_pos[:1], o_pos[1:].unsqueeze(0), 0.3, 1.4)
    # identical uint8 network inputs (bit-exact resize) -> the scores agree to the run-to-run noise of the fp32 atomics
    assert (got["output_blocks"][0].cpu() - o[1:]).abs().max().item() <= 1e-5
    assert (got["output_final"].cpu() - want).abs().max().item() <= 1e-4
    assert (got["output_pos_final"].cpu() - want_pos).abs().max().item() <= 1e-4


# ---------------------------------------------------------------------------------------------------------------------
# the adapter variant of the model (trainers/Caption_distill_double_adapter.py)
# ---------------------------------------------------------------------------------------------------------------------
def test_adapter_model_matches_reference():
    """`AdapterDenseCLIPB200` vs the reference `AdapterDenseCLIP`: state_dict keys, test 4-tuple, train 4-tuple, ranking loss and
    the gradients of ctx / ctx_double through the frozen residual adapter."""
    from lecb200.adapter_clip import AdapterDenseCLIPB200
    from lecb200.clip_model import CLIPParams
    from oracle.ref_extract import make_cfg
    from .test_oracle_ext import adapter_case
    L = _losses()
    c = adapter_case()
    g = c["gold"]
    arch = c["arch"]
    clip = CLIPParams(*arch.ctor_args())
    clip.load_state_dict(c["sd"], strict=False)
    clip = clip.float().cuda().eval()
    model = AdapterDenseCLIPB200(make_cfg(arch.image_resolution, n_ctx=c["n_ctx"]), c["names"], clip, tokenized_prompts=c["tokens"]).cuda()
    assert sorted(model.state_dict().keys()) == [str(k) for k in g["state_keys"]]
    wd, wu = c["adapter"]
    with torch.no_grad():
        model.prompt_learner.ctx.copy_(c["pl_state"]["ctx"])
        model.prompt_learner.ctx_double.copy_(c["pl_state"]["ctx_double"])
        model.adapter_text_encoder.text_adapter.fc[0].weight.copy_(wd)
        model.adapter_text_encoder.text_adapter.fc[2].weight.copy_(wu)
    model.adapter_text_encoder.refresh_adapter()
    for name, p in model.named_parameters():
        if "prompt_learner" not in name:
            p.requires_grad_(False)          # TA:534-536
    out = model(c["image"].cuda(), if_test=True)
    assert len(out) == 4
    errs = [np.abs(t.float().cpu().numpy() - g["test_" + n]).max() for t, n in zip(out, ("logits", "logits_local", "neg_map", "pos_map"))]
    print(f"[adapter] test errors {errs}")
    assert max(errs) <= LOGIT_TOL
    r = model(None, c["captions"].cuda())
    assert len(r) == 4
    y = c["labels"].cuda()
    loss = L.ranking_loss(r[0], y, scale_=1.0, margin_=1) + L.ranking_loss(r[1], y, scale_=1.0, margin_=1)
    loss.backward()
    torch.cuda.synchronize()
    e1 = np.abs(r[0].detach().cpu().numpy() - g["train_logits"]).max()
    e2 = np.abs(r[1].detach().cpu().numpy() - g["train_logits_local"]).max()
    print(f"[adapter] train logits err {e1:.5f} local err {e2:.5f} loss {loss.item():.4f} vs {float(g['loss']):.4f}")
    assert e1 <= LOGIT_TOL and e2 <= LOGIT_TOL
    assert abs(loss.item() - float(g["loss"])) <= 1e-2 * max(1.0, abs(float(g["loss"])))
    np.testing.assert_allclose(r[3].detach().cpu().numpy(), g["train_text_features"], atol=5e-3)
    for pname in ("ctx", "ctx_double"):
        gref = g["grad_" + pname]
        got = getattr(model.prompt_learner, pname).grad.cpu().numpy()
        err = np.abs(got - gref).max() / np.abs(gref).max()
        cos = float((got.flatten() @ gref.flatten()) / (np.linalg.norm(got) * np.linalg.norm(gref)))
        print(f"[adapter] grad {pname}: max err / max ref = {err:.4f}, cosine = {cos:.5f}")
        # The two ReLUs of the adapter make this gradient discontinuous in the adapter's input: the REFERENCE's own fp32
        # gradient moves by 5-11 % of its max (cosine 0.995-0.999) when that input is perturbed by 1e-3 relative or the
        # adapter's operands are rounded to bf16, and by 14-18 % (cosine 0.987-0.990) at 5e-3 — the bf16 transformer's
        # error band (tools/adapter_grad_sensitivity.py).  The backward kernels themselves are pinned to 1e-2 with the
        # masks held fixed in test_adapter_backward_chain_matches_fp32.
        assert err < 0.2 and cos > 0.985, (pname, err, cos)
    for name, p in model.named_parameters():
        if "prompt_learner" not in name:
            assert p.grad is None, name


@pytest.mark.gpu
def test_adapter_backward_chain_matches_fp32():
    """TextTower._adapter_fwd / _adapter_bwd against fp32 torch with the ReLU masks taken from the kernel's own forward
    (a1 > 0, z2 > 0), so that a mask flip at a rounding-level pre-activation cannot hide or fake a backward error."""
    from lecb200 import synth
    from lecb200.engine import TextTower
    tower = TextTower.__new__(TextTower)
    tower.device = torch.device("cuda")
    wd, wu = synth.adapter_weights(5)
    tower.set_adapter(wd, wu)
    gen = torch.Generator().manual_seed(11)
    xr = torch.randn((160, 512), generator=gen).cuda()
    d_out = torch.randn((160, 512), generator=gen).cuda()
    out, a1, z2 = tower._adapter_fwd(xr)
    wdf, wuf = wd.cuda().bfloat16().float(), wu.cuda().bfloat16().float()
    xb = xr.bfloat16().float()
    a1_ref = torch.relu(xb @ wdf.t())
    ref = xr + torch.relu(a1_ref.bfloat16().float() @ wuf.t())
    assert (out - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert (a1.float() - a1_ref).abs().max().item() <= 2e-2 * a1_ref.abs().max().item()
    got = tower._adapter_bwd(d_out, a1, z2)
    m2, m1 = (z2 > 0).float(), (a1.float() > 0).float()
    want = d_out + (((d_out * m2) @ wuf) * m1) @ wdf
    err = (got - want).abs().max().item() / want.abs().max().item()
    print(f"[adapter bwd chain] max err / max = {err:.5f}")
    assert err <= 1e-2
