"""Live pin: run the reference's OWN classes (AST-extracted from /root/reference) next to oracle/restatement.py.
Only runs in the build container (the reference tree does not exist on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_extract as RX
from oracle import restatement as R
from oracle import synth

pytestmark = pytest.mark.skipif(not RX.available(), reason="/root/reference not present")


def _tiny_setup(use_evidence, bank_rows=48, seed=321):
    from oracle.make_golden import TINY_CLASSES, build_dense_clip
    arch = synth.tiny_rn()
    sd = synth.clip_state_dict(arch, 0)
    bank = synth.caption_bank(bank_rows, arch.embed_dim, seed)
    model = build_dense_clip(arch, sd, TINY_CLASSES, 4, use_evidence, bank, seed)
    toks = model.tokenized_prompts
    ctx = [synth.prompt_ctx(4, arch.transformer_width, seed, t) for t in ("pos", "neg", "evi")]
    return arch, sd, bank, model, toks, R.prompt_learner_state(sd, toks, 4, *ctx)


@pytest.mark.parametrize("ev", [False, True])
def test_inference_path_live(ev):
    arch, sd, bank, model, toks, pl = _tiny_setup(ev)
    img = synth.images(2, arch.image_resolution, 5)
    with torch.no_grad():
        ref = model(img, if_test=True)
        got = R.dense_clip_test(sd, arch, img, pl, toks, use_evidence=ev, bank=bank)
    for a, b in zip(ref, got):
        np.testing.assert_allclose(b.numpy(), a.float().numpy(), atol=2e-5, rtol=1e-4)


def test_train_path_and_losses_live():
    arch, sd, bank, model, toks, pl = _tiny_setup(True)
    L = RX.loss_functions()
    caps = synth.captions(3, 9, vocab=arch.vocab_size)
    y = synth.labels(3, toks.shape[0], 9)
    ref = model(None, caps)
    loss_ref = L["ASL_loss"](ref[0], y) + L["ranking_loss"](ref[1] * 1.0, y, scale_=1.0, margin_=1)
    loss_ref.backward()
    pl = {k: (v.clone().requires_grad_(True) if k.startswith("ctx") else v) for k, v in pl.items()}
    got = R.dense_clip_train(sd, arch, caps, pl, toks, use_evidence=True)
    loss = R.asl_loss(got[0], y) + R.ranking_loss(got[1], y, 1.0, 1.0)
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
    for name in ("ctx", "ctx_double", "ctx_evidence"):
        g_ref = getattr(model.prompt_learner, name).grad
        scale = g_ref.abs().max().item() + 1e-12
        np.testing.assert_allclose(pl[name].grad.numpy() / scale, g_ref.numpy() / scale, atol=2e-4)


def test_map_live():
    ref_map = RX.mAP_function()
    rng = np.random.default_rng(0)
    s, t = rng.standard_normal((40, 7)), (rng.random((40, 7)) < 0.2).astype(np.float32)
    t[0] = 1
    assert abs(ref_map(t, s) - R.mean_average_precision(t, s)) < 1e-9
