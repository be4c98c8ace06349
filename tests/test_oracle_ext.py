"""Round-2 switches of the prompt-tuning step: oracle/restatement.py against the reference-generated fixtures
train_ext_*.npz / losses_ext.npz / prompt_learner_tiny.npz (CPU; no GPU, no reference tree).

Covers TRAIN.ema (T:516-541, 554-559, 809-813), TRAINER.Caption.CSC (T:127-133), TRAIN.IF_LEARN_SCALE (T:453-454, 493-494),
LOSSFUNC 'ranking_with_cooccurrence' (T:842-850, U:95-110), PromptLearner.forward(neg_prompt_wcls=False) and name_lens."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import synth

from . import _cases as C

ATOL = 2e-4


def ext_case(tag):
    g = C.load(f"train_ext_{tag}.npz")
    arch = C.TRAIN_CASES[tag]()
    seed = int(g["seed"])
    sd = synth.clip_state_dict(arch, 0)
    toks, n_ctx, names = C.tokens_for("tiny" if tag == "tiny" else "coco")
    caps = synth.captions(int(g["batch"]), seed, vocab=arch.vocab_size)
    y = synth.labels(int(g["batch"]), len(names), seed)
    np.testing.assert_allclose(C.checksum(caps.float()), g["caption_checksum"], rtol=1e-12, err_msg="RNG drift: captions")
    np.testing.assert_allclose(C.checksum(y), g["label_checksum"], rtol=1e-12, err_msg="RNG drift: labels")
    return dict(arch=arch, sd=sd, captions=caps, labels=y, tokens=toks, n_ctx=n_ctx, names=names, gold=g, seed=seed)


def pl_state(c, csc=False, twin=False):
    """Prompt-learner state of the fixture: contexts from the seeded generator (oracle/make_golden.build_dense_clip)."""
    w = c["arch"].transformer_width
    n_cls = len(c["names"]) if csc else 0
    ctx = synth.prompt_ctx(c["n_ctx"], w, c["seed"], "pos", n_cls)
    ctx_d = synth.prompt_ctx(c["n_ctx"], w, c["seed"], "neg", n_cls)
    ctx_e = synth.prompt_ctx(c["n_ctx"], w, c["seed"], "evi")
    if twin:
        ctx = ctx + synth.prompt_ctx(c["n_ctx"], w, c["seed"], "pos_m", n_cls)
        ctx_d = ctx_d + synth.prompt_ctx(c["n_ctx"], w, c["seed"], "neg_m", n_cls)
        ctx_e = ctx_e + synth.prompt_ctx(c["n_ctx"], w, c["seed"], "evi_m")
    return R.prompt_learner_state(c["sd"], c["tokens"], c["n_ctx"], ctx, ctx_d, ctx_e)


def _grads_close(pl, g, sfx, names=("ctx", "ctx_double", "ctx_evidence"), atol=2e-4):
    from oracle.make_golden import CSC_ROWS
    for pname in names:
        gref = g[f"grad_{pname}" + sfx]
        got = pl[pname].grad
        if bool(g[f"gradnone_{pname}" + sfx]):
            assert got is None or float(got.abs().max()) == 0.0, pname
            continue
        got = got.numpy()
        if got.ndim == 3 and got.shape[0] != gref.shape[0]:
            np.testing.assert_allclose(np.linalg.norm(got.reshape(got.shape[0], -1), axis=1), g[f"gradnorm_{pname}" + sfx],
                                       rtol=2e-3, atol=1e-6)
            got = got[list(CSC_ROWS)]
        scale = max(np.abs(gref).max(), 1e-8)
        np.testing.assert_allclose(got / scale, gref / scale, atol=atol, err_msg=f"{sfx}:{pname}")


def _rank2(out, y):
    return R.ranking_loss(out[0], y, 1.0, 1.0) + R.ranking_loss(out[1], y, 1.0, 1.0)


@pytest.mark.parametrize("tag", ["tiny", "rn50"])
@pytest.mark.parametrize("ev", [False, True])
def test_ema_matches_reference(tag, ev):
    c = ext_case(tag)
    g = c["gold"]
    sfx = "_ema" + ("_ev" if ev else "")
    pl = {k: (v.clone().requires_grad_(True) if k.startswith("ctx") else v) for k, v in pl_state(c).items()}
    out = R.dense_clip_train_ema(c["sd"], c["arch"], c["captions"], pl, pl_state(c, twin=True), c["tokens"], use_evidence=ev)
    r_loss = _rank2(out, c["labels"])
    e_loss = R.ema_loss(out[0], out[4], out[1], out[5])
    assert abs(r_loss.item() - float(g["r_loss" + sfx])) < 1e-4 * max(1.0, abs(float(g["r_loss" + sfx])))
    assert abs(e_loss.item() - float(g["ema_loss" + sfx])) < 2e-3 * max(1.0, abs(float(g["ema_loss" + sfx])))
    (r_loss + e_loss).backward()
    np.testing.assert_allclose(out[4].numpy(), g["logits_m" + sfx], atol=ATOL)
    np.testing.assert_allclose(out[5].numpy(), g["logits_local_m" + sfx], atol=ATOL)
    for pname in ("ctx", "ctx_double", "ctx_evidence"):
        np.testing.assert_allclose(out[6][pname].numpy(), g[f"twin_{pname}" + sfx], atol=1e-7)
    # the 10000 x local KL term dominates the gradient and amplifies fp32 op-order noise: 2e-3 of the max
    _grads_close(pl, g, sfx, atol=2e-3)


@pytest.mark.parametrize("tag", ["tiny", "rn50"])
@pytest.mark.parametrize("ev", [False, True])
def test_learnable_scale_matches_reference(tag, ev):
    c = ext_case(tag)
    g = c["gold"]
    sfx = "_scale" + ("_ev" if ev else "")
    pl = {k: (v.clone().requires_grad_(True) if k.startswith("ctx") else v) for k, v in pl_state(c).items()}
    temperature = torch.tensor(3.0, requires_grad=True)                     # PromptLearner init, T:160-161
    out = R.dense_clip_train(c["sd"], c["arch"], c["captions"], pl, c["tokens"], use_evidence=ev, logit_scale=temperature.exp())
    loss = _rank2(out, c["labels"])
    loss.backward()
    assert abs(loss.item() - float(g["loss" + sfx])) < 1e-4 * max(1.0, abs(float(g["loss" + sfx])))
    np.testing.assert_allclose(out[0].detach().numpy(), g["logits" + sfx], atol=5 * ATOL)          # logits x exp(3) = 20
    np.testing.assert_allclose(out[1].detach().numpy(), g["logits_local" + sfx], atol=5 * ATOL)
    gt = float(g["grad_temperature" + sfx])
    # d loss / d temperature = sum(dlogits * logits): a small residual of large cancelling terms (fp32 op-order noise)
    assert abs(temperature.grad.item() - gt) < 2e-3 * max(1.0, abs(gt))
    _grads_close(pl, g, sfx)


@pytest.mark.parametrize("tag", ["tiny", "rn50"])
def test_csc_matches_reference(tag):
    c = ext_case(tag)
    g = c["gold"]
    pl = {k: (v.clone().requires_grad_(True) if k.startswith("ctx") else v) for k, v in pl_state(c, csc=True).items()}
    out = R.dense_clip_train(c["sd"], c["arch"], c["captions"], pl, c["tokens"], use_evidence=True)
    loss = _rank2(out, c["labels"])
    loss.backward()
    assert abs(loss.item() - float(g["loss_csc_ev"])) < 1e-4 * max(1.0, abs(float(g["loss_csc_ev"])))
    np.testing.assert_allclose(out[3].detach().numpy(), g["text_features_csc_ev"], atol=1e-5)
    _grads_close(pl, g, "_csc_ev")


def test_cooccurrence_ranking_matches_reference():
    c = ext_case("rn50")
    g = c["gold"]
    p = torch.from_numpy(g["cooc_p"])
    pl = {k: (v.clone().requires_grad_(True) if k.startswith("ctx") else v) for k, v in pl_state(c).items()}
    out = R.dense_clip_train(c["sd"], c["arch"], c["captions"], pl, c["tokens"], use_evidence=False)
    loss = R.ranking_loss_with_cooccurrence(out[0], c["labels"], p, 1.0, 1.0) + \
        R.ranking_loss_with_cooccurrence(out[1], c["labels"], p, 1.0, 1.0)
    loss.backward()
    assert abs(loss.item() - float(g["loss_cooc"])) < 1e-4 * max(1.0, abs(float(g["loss_cooc"])))
    _grads_close(pl, g, "_cooc")


def test_losses_ext_match_reference():
    g = C.load("losses_ext.npz")
    x0, xm, y, p = (torch.from_numpy(g[k]) for k in ("x", "xm", "y", "cooc_p"))
    for name, fn in (("cooc_s1", lambda a: R.ranking_loss_with_cooccurrence(a, y, p, 1.0, 1.0)),
                     ("cooc_s2", lambda a: R.ranking_loss_with_cooccurrence(a, y, p)),
                     ("kl", lambda a: R.kl_softmax(a, xm)),
                     ("kl_x10000", lambda a: R.kl_softmax(a, xm) * 10000)):
        a = x0.clone().requires_grad_(True)
        loss = fn(a)
        loss.backward()
        ref = float(g["loss_" + name])
        assert abs(loss.item() - ref) < 1e-5 * max(1.0, abs(ref)), name
        np.testing.assert_allclose(a.grad.numpy(), g["grad_" + name], atol=1e-6 * max(1.0, np.abs(g["grad_" + name]).max()),
                                   rtol=1e-4, err_msg=name)


@pytest.mark.parametrize("csc", [False, True])
def test_prompt_learner_mirror_matches_reference(csc):
    """lecb200's PromptLearner (pure torch plumbing, the one module of the mirror that runs on the CPU) against the
    reference class: prompts for neg_prompt_wcls True / False, name_lens, tokenized prompts and every state_dict entry."""
    pytest.importorskip("lecb200")
    from lecb200.clip_model import CLIPParams
    from lecb200.dense_clip import PromptLearner
    from oracle.ref_extract import make_cfg
    g = C.load("prompt_learner_tiny.npz")
    sfx = "_csc" if csc else ""
    arch = synth.tiny_rn()
    clip = CLIPParams(*arch.ctor_args())
    clip.load_state_dict(synth.clip_state_dict(arch, 0), strict=False)
    clip = clip.float().eval()
    toks, n_ctx, names = C.tokens_for("tiny")
    pl = PromptLearner(make_cfg(arch.image_resolution, n_ctx=4, csc=csc), names, clip, tokenized_prompts=toks)
    n_cls = len(names) if csc else 0
    with torch.no_grad():
        pl.ctx.copy_(synth.prompt_ctx(4, arch.transformer_width, 7, "pos", n_cls))
        pl.ctx_double.copy_(synth.prompt_ctx(4, arch.transformer_width, 7, "neg", n_cls))
        pl.ctx_evidence.copy_(synth.prompt_ctx(4, arch.transformer_width, 7, "evi"))
    assert list(pl.name_lens) == list(g["name_lens" + sfx])
    np.testing.assert_array_equal(pl.tokenized_prompts.numpy(), g["tokenized_prompts" + sfx])
    for wcls in (True, False):
        r = pl(neg_prompt_wcls=wcls)
        for nm, t in zip(("prompts", "prompts_neg", "prompts_evidence"), r[:3]):
            np.testing.assert_array_equal(t.detach().numpy(), g[f"{nm}_wcls{int(wcls)}{sfx}"], err_msg=f"{nm} wcls={wcls}")
    sd = pl.state_dict()
    want = {k[len("state_"):-len(sfx)] if sfx else k[len("state_"):] for k in g.files if k.startswith("state_") and k.endswith(sfx)
            and (sfx or not k.endswith("_csc"))}
    assert set(sd) == want, set(sd) ^ want
    for k, v in sd.items():
        np.testing.assert_array_equal(v.numpy(), g[f"state_{k}{sfx}"], err_msg=k)


def test_resample_loss_matches_reference():
    """`ResampleLoss` (trainers/dbl.py, LOSSFUNC 'dbl'): restatement vs the reference class on the three fixture settings."""
    from oracle.make_golden import RESAMPLE_CONFIGS
    g = C.load("resample_loss.npz")
    x0, y = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    cf, ncf = torch.from_numpy(g["class_freq"]), torch.from_numpy(g["neg_class_freq"])
    for name, kw in RESAMPLE_CONFIGS.items():
        a = x0.clone().requires_grad_(True)
        loss = R.resample_loss(a, y, cf, ncf, reweight=kw["reweight_func"] == "rebalance", map_alpha=kw["map_param"]["alpha"],
                               map_beta=kw["map_param"]["beta"], map_gamma=kw["map_param"]["gamma"], logit_reg=kw["logit_reg"],
                               focal=kw["focal"]["focal"], focal_gamma=kw["focal"]["gamma"], balance_param=kw["focal"]["balance_param"],
                               loss_weight=kw["loss_weight"])
        loss.backward()
        ref = float(g["loss_" + name])
        assert abs(loss.item() - ref) < 1e-5 * max(1.0, abs(ref)), name
        np.testing.assert_allclose(a.grad.numpy(), g["grad_" + name], atol=1e-7, rtol=1e-4, err_msg=name)
        np.testing.assert_array_equal(a.detach().numpy(), g["x"])


def adapter_case():
    g = C.load("adapter_rn50.npz")
    arch = synth.RN50(224)
    seed = int(g["seed"])
    sd = synth.clip_state_dict(arch, 0)
    toks, n_ctx, names = C.tokens_for("coco")
    img = synth.images(int(g["batch_img"]), arch.image_resolution, seed)
    caps = synth.captions(int(g["batch_cap"]), seed, vocab=arch.vocab_size)
    y = synth.labels(int(g["batch_cap"]), len(names), seed)
    np.testing.assert_allclose(C.checksum(img), g["image_checksum"], rtol=1e-12, err_msg="RNG drift: images")
    np.testing.assert_allclose(C.checksum(caps.float()), g["caption_checksum"], rtol=1e-12, err_msg="RNG drift: captions")
    w = arch.transformer_width
    pl = R.prompt_learner_state(sd, toks, n_ctx, synth.prompt_ctx(n_ctx, w, seed, "pos"), synth.prompt_ctx(n_ctx, w, seed, "neg"),
                                synth.prompt_ctx(n_ctx, w, seed, "evi"))
    return dict(arch=arch, sd=sd, image=img, captions=caps, labels=y, tokens=toks, n_ctx=n_ctx, names=names, gold=g, seed=seed,
                pl_state=pl, adapter=synth.adapter_weights(seed))


def test_adapter_model_matches_reference():
    """`AdapterDenseCLIP` (trainers/Caption_distill_double_adapter.py:320-457): restatement vs the reference class."""
    c = adapter_case()
    g = c["gold"]
    wd, wu = c["adapter"]
    with torch.no_grad():
        out = R.adapter_dense_clip_test(c["sd"], c["arch"], c["image"], c["pl_state"], c["tokens"], wd, wu)
    for name, t in zip(("logits", "logits_local", "neg_map", "pos_map"), out):
        np.testing.assert_allclose(t.numpy(), g["test_" + name], atol=ATOL, rtol=1e-4, err_msg=name)
    pl = {k: (v.clone().requires_grad_(True) if k in ("ctx", "ctx_double") else v) for k, v in c["pl_state"].items()}
    r = R.adapter_dense_clip_train(c["sd"], c["arch"], c["captions"], pl, c["tokens"], wd, wu)
    loss = R.ranking_loss(r[0], c["labels"], 1.0, 1.0) + R.ranking_loss(r[1], c["labels"], 1.0, 1.0)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    np.testing.assert_allclose(r[0].detach().numpy(), g["train_logits"], atol=ATOL)
    np.testing.assert_allclose(r[1].detach().numpy(), g["train_logits_local"], atol=ATOL)
    np.testing.assert_allclose(r[3].detach().numpy(), g["train_text_features"], atol=1e-5)
    for pname in ("ctx", "ctx_double"):
        gref = g["grad_" + pname]
        scale = np.abs(gref).max()
        np.testing.assert_allclose(pl[pname].grad.numpy() / scale, gref / scale, atol=2e-4, err_msg=pname)
