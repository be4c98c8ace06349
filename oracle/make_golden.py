"""Generate tests/golden/* by RUNNING THE REFERENCE'S OWN CLASSES (build container only).

    python -m oracle.make_golden            # needs /root/reference; ~2-3 min on 8 CPU threads

The reference ships no tests, golden vectors or fixtures for this path, so these
reference-generated outputs are what pins both `oracle/restatement.py` and the CUDA path.  Inputs and
weights are regenerated from seeds (oracle/synth.py) on the consuming side; each fixture stores
checksums of them so RNG drift is caught instead of silently comparing different inputs.

Fixtures written:
  prompt_tokens_coco80.npz   reference tokenizer output for "X X ... X <class>." (n_ctx 16) + class names
  head_rn50_224.npz          cfg 1: DenseCLIP.forward(image, if_test=True), RN50 224² B=8, evidence off/on
  head_rn101_448.npz         cfg 2 shape at B=2 (deviation 1: T:447 literal 1024 -> 512)
  head_tiny.npz              toy ModifiedResNet (width 8, 64² images) for fast CPU tests
  head_small.npz             width-64, one-block-per-stage ModifiedResNet at 128² (GPU kernel tests)
  train_rn50.npz             DenseCLIP.forward(None, captions) B=4 + ranking/ASL losses + prompt grads
  train_tiny.npz             same on the toy arch, B=6
  losses.npz                 ranking_loss / ASL_loss / dualcoop_loss values + grads on [16,80]
  map.npz                    reference numpy mAP on synthetic scores/labels
  fusion.npz                 reference fuse / fuse6 (gen_final_ans.py) and adjust_predictions (T:611-615) on synthetic scores
  train_ext_{tiny,rn50}.npz  the same step under TRAIN.ema / CSC / IF_LEARN_SCALE / co-occurrence ranking (round 2)
  losses_ext.npz             ranking_loss_with_cooccurrence + the KL terms of T:809-813 on [16,80]
  resample_loss.npz          ResampleLoss (trainers/dbl.py, LOSSFUNC 'dbl'): shipped, focal + logit_reg and unweighted settings
  adapter_rn50.npz           AdapterDenseCLIP (trainers/Caption_distill_double_adapter.py): test / train tuples, loss, prompt gradients
  prompt_learner_tiny.npz    PromptLearner.forward(neg_prompt_wcls=True/False), CSC and generic, name_lens, state_dict
  vit_{tiny,b16_224,l14_224}.npz   reference VisionTransformer.forward (class-token feature) on synthetic weights
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from . import ref_extract as RX
from . import synth

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CSC_ROWS = (0, 7, 33, 52, 79)          # classes whose CSC gradients are stored in full (train_ext_rn50.npz)
TINY_CLASSES = ["person", "dog", "traffic light", "cup", "potted plant", "tv"]


def checksum(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64) % 97 + 1
    return np.array([t.sum().item(), (t * w).sum().item(), t.abs().max().item()], dtype=np.float64)


def state_checksum(sd) -> np.ndarray:
    acc = np.zeros(3)
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            acc += checksum(v)
    return acc


def build_dense_clip(arch, sd, classnames, n_ctx, use_evidence, bank, seed, csc=False, ema=False, learn_scale=False,
                     twin_offset=False):
    ns = RX.trainer_classes(bank, arch.embed_dim)
    clip_model = RX.build_reference_clip(arch, sd)
    cfg = RX.make_cfg(arch.image_resolution, n_ctx=n_ctx, csc=csc, use_evidence=use_evidence, ema=ema,
                      learn_scale=learn_scale)
    model = ns["DenseCLIP"](cfg, classnames, clip_model)
    w = arch.transformer_width
    n_cls = len(classnames) if csc else 0
    with torch.no_grad():
        for name, tag in (("ctx", "pos"), ("ctx_double", "neg")):
            getattr(model.prompt_learner, name).copy_(synth.prompt_ctx(n_ctx, w, seed, tag, n_cls))
        model.prompt_learner.ctx_evidence.copy_(synth.prompt_ctx(n_ctx, w, seed, "evi"))          # never CSC (T:147)
    model.copy_params()
    if twin_offset:
        # copy_params() makes the twin equal to the live learner, which would make the momentum update (T:554-559) a no-op:
        # start the twin somewhere else so that `0.995 twin + 0.005 live` and the twin's logits are observable
        with torch.no_grad():
            for name, tag in (("ctx", "pos_m"), ("ctx_double", "neg_m")):
                getattr(model.prompt_learner_m, name).add_(synth.prompt_ctx(n_ctx, w, seed, tag, n_cls))
            model.prompt_learner_m.ctx_evidence.add_(synth.prompt_ctx(n_ctx, w, seed, "evi_m"))
    for name, p in model.named_parameters():
        if "prompt_learner" not in name:
            p.requires_grad_(False)          # T:763-765
    return model


def golden_tokens(classnames):
    clip_pkg = RX.clip_package()
    prefix = " ".join(["X"] * 16)
    toks = torch.cat([clip_pkg.clip.tokenize(prefix + " " + n.replace("_", " ") + ".", truncate=True) for n in classnames])
    toks_nocls = torch.cat([clip_pkg.clip.tokenize(prefix + ".", truncate=True) for _ in classnames])
    tiny = torch.cat([clip_pkg.clip.tokenize(" ".join(["X"] * 4) + " " + n + ".", truncate=True) for n in TINY_CLASSES])
    np.savez_compressed(os.path.join(GOLD, "prompt_tokens_coco80.npz"),
                        tokens=toks.numpy(), tokens_nocls=toks_nocls.numpy(), n_ctx=np.int64(16),
                        classnames=np.array(classnames), tiny_tokens=tiny.numpy(),
                        tiny_classnames=np.array(TINY_CLASSES), tiny_n_ctx=np.int64(4))
    return toks, tiny


def golden_head(tag, arch, batch, classnames, n_ctx, bank_rows, seed, evidence_modes=(False, True)):
    sd = synth.clip_state_dict(arch, seed=0)
    img = synth.images(batch, arch.image_resolution, seed)
    bank = synth.caption_bank(bank_rows, arch.embed_dim, seed)
    out = {"image_checksum": checksum(img), "state_checksum": state_checksum(sd),
           "bank_checksum": checksum(bank.float()), "batch": np.int64(batch), "seed": np.int64(seed),
           "bank_rows": np.int64(bank_rows), "n_ctx": np.int64(n_ctx)}
    for ev in evidence_modes:
        t0 = time.time()
        model = build_dense_clip(arch, sd, classnames, n_ctx, ev, bank, seed)
        with torch.no_grad():
            r = model(img, if_test=True)
        sfx = "_ev" if ev else ""
        for name, t in zip(("logits", "logits_local", "neg_map", "pos_map", "topk_scores"), r):
            out[name + sfx] = t.float().numpy()
        # cached, normalised prompt features (T:421-439): handy intermediate for kernel debugging
        for k, v in model.prompt_text_features.items():
            out[k + sfx] = v.float().numpy()
        print(f"[{tag}] evidence={ev}: {time.time() - t0:.1f}s  logits absmax={r[0].abs().max():.4f} "
              f"local absmax={r[1].abs().max():.4f}", flush=True)
    np.savez_compressed(os.path.join(GOLD, f"head_{tag}.npz"), **out)


def golden_train(tag, arch, batch, classnames, n_ctx, seed):
    sd = synth.clip_state_dict(arch, seed=0)
    caps = synth.captions(batch, seed, vocab=arch.vocab_size)
    y = synth.labels(batch, len(classnames), seed)
    L = RX.loss_functions()
    out = {"caption_checksum": checksum(caps.float()), "state_checksum": state_checksum(sd),
           "label_checksum": checksum(y), "batch": np.int64(batch), "seed": np.int64(seed), "n_ctx": np.int64(n_ctx)}
    for ev in (False, True):
        for loss_name in ("ranking", "asl"):
            model = build_dense_clip(arch, sd, classnames, n_ctx, ev, None, seed)
            r = model(None, caps)
            logits, logits_local = r[0], r[1]
            if loss_name == "ranking":       # T:806-808
                loss = L["ranking_loss"](logits, y, scale_=1.0, margin_=1) + \
                       L["ranking_loss"](logits_local, y, scale_=1.0, margin_=1)
            else:
                loss = L["ASL_loss"](logits, y) + L["ASL_loss"](logits_local, y)
            loss.backward()
            sfx = ("_ev" if ev else "") + "_" + loss_name
            out["loss" + sfx] = np.float64(loss.item())
            pl = model.prompt_learner
            for pname in ("ctx", "ctx_double", "ctx_evidence"):
                g = getattr(pl, pname).grad
                out[f"grad_{pname}" + sfx] = (torch.zeros_like(getattr(pl, pname)) if g is None else g).numpy()
                out[f"gradnone_{pname}" + sfx] = np.bool_(g is None)
            if loss_name == "ranking":
                s2 = "_ev" if ev else ""
                out["logits" + s2] = logits.detach().numpy()
                out["logits_local" + s2] = logits_local.detach().numpy()
                out["seq_feats" + s2] = r[2].detach().numpy()[:, :2]      # [L,2,D] slice keeps the file small
                out["text_features" + s2] = r[3].detach().numpy()
            print(f"[train_{tag}] evidence={ev} loss={loss_name}: {loss.item():.6f}", flush=True)
    np.savez_compressed(os.path.join(GOLD, f"train_{tag}.npz"), **out)


def _cooc_p():
    """The co-occurrence prior the trainer builds from freq_stats.pkl (T:843-846): an INPUT of the loss, stored in the fixture."""
    import pickle
    with open(os.path.join(RX.MC, "freq_stats.pkl"), "rb") as f:
        result = pickle.load(f)
    p = torch.tensor(result["adj"] / result["nums"][:, np.newaxis], dtype=torch.float32)
    return p / p.sum(-1)[:, None]


def golden_train_ext(tag, arch, batch, classnames, n_ctx, seed):
    """Config switches of the prompt-tuning step that train_<tag>.npz does not cover, each run through the reference's own
    DenseCLIP.forward(None, captions) and the loss expression of forward_backward (T:804-815, T:842-850):
      ema      TRAIN.ema=True: 6-tuple with logits_m_ / logits_local_m after _momentum_update (T:516-541, 554-559), loss =
               ranking + KL(log_softmax(out) || softmax(out_m)) + 10000 x the local KL (T:809-813); twin parameters after the update
      csc      TRAINER.Caption.CSC=True: class-specific ctx / ctx_double [K,n_ctx,W] (T:127-133)
      scale    TRAIN.IF_LEARN_SCALE=True: logit scale exp(temperature) with its gradient (T:453-454, 493-494)
      cooc     LOSSFUNC 'ranking_with_cooccurrence' (T:842-850, U:95-110) on the plain configuration"""
    sd = synth.clip_state_dict(arch, seed=0)
    caps = synth.captions(batch, seed, vocab=arch.vocab_size)
    y = synth.labels(batch, len(classnames), seed)
    L = RX.loss_functions()
    kl = torch.nn.KLDivLoss(reduction="batchmean")
    F = torch.nn.functional
    out = {"caption_checksum": checksum(caps.float()), "state_checksum": state_checksum(sd),
           "label_checksum": checksum(y), "batch": np.int64(batch), "seed": np.int64(seed), "n_ctx": np.int64(n_ctx)}
    if len(classnames) == 80:
        out["cooc_p"] = _cooc_p().numpy()

    def record(sfx, model, r, loss):
        loss.backward()
        out["loss" + sfx] = np.float64(loss.item())
        pl = model.prompt_learner
        for pname in ("ctx", "ctx_double", "ctx_evidence", "temperature"):
            g = getattr(pl, pname).grad
            gt = torch.zeros_like(getattr(pl, pname)) if g is None else g
            if gt.dim() == 3 and gt.shape[0] > len(CSC_ROWS):
                # class-specific contexts [K,n_ctx,W]: keep a few classes in full and every class's norm (file size)
                out[f"gradnorm_{pname}" + sfx] = gt.flatten(1).norm(dim=1).numpy()
                gt = gt[list(CSC_ROWS)]
            out[f"grad_{pname}" + sfx] = gt.numpy()
            out[f"gradnone_{pname}" + sfx] = np.bool_(g is None)
        out["logits" + sfx] = r[0].detach().numpy()
        out["logits_local" + sfx] = r[1].detach().numpy()
        out["text_features" + sfx] = r[3].detach().numpy()
        print(f"[train_ext_{tag}] {sfx}: loss {loss.item():.6f}", flush=True)

    def rank2(a, b):
        return L["ranking_loss"](a, y, scale_=1.0, margin_=1) + L["ranking_loss"](b, y, scale_=1.0, margin_=1)

    for ev in (False, True):
        e = "_ev" if ev else ""
        # --- ema ---
        model = build_dense_clip(arch, sd, classnames, n_ctx, ev, None, seed, ema=True, twin_offset=True)
        r = model(None, caps)
        r_loss = rank2(r[0], r[1])
        # ranking_loss scaled its arguments in place by 1.0 (U:86): values unchanged
        ema_loss = kl(F.log_softmax(r[0], dim=-1), F.softmax(r[4], dim=-1)) + \
            kl(F.log_softmax(r[1], dim=-1), F.softmax(r[5], dim=-1)) * 10000
        out["r_loss_ema" + e] = np.float64(r_loss.item())
        out["ema_loss_ema" + e] = np.float64(ema_loss.item())
        out["logits_m_ema" + e] = r[4].detach().numpy()
        out["logits_local_m_ema" + e] = r[5].detach().numpy()
        for pname in ("ctx", "ctx_double", "ctx_evidence"):
            out[f"twin_{pname}_ema" + e] = getattr(model.prompt_learner_m, pname).detach().numpy().copy()
        record("_ema" + e, model, r, r_loss + ema_loss)
        # --- learnable logit scale ---
        model = build_dense_clip(arch, sd, classnames, n_ctx, ev, None, seed, learn_scale=True)
        r = model(None, caps)
        record("_scale" + e, model, r, rank2(r[0], r[1]))
    # --- class-specific contexts (no evidence: T:147 keeps ctx_evidence generic either way) ---
    model = build_dense_clip(arch, sd, classnames, n_ctx, True, None, seed, csc=True)
    r = model(None, caps)
    record("_csc_ev", model, r, rank2(r[0], r[1]))
    # --- co-occurrence weighted ranking loss ---
    if len(classnames) == 80:
        p = _cooc_p()
        model = build_dense_clip(arch, sd, classnames, n_ctx, False, None, seed)
        r = model(None, caps)
        loss = L["ranking_loss_with_cooccurrence"](r[0], y, p, scale_=1.0, margin_=1) + \
            L["ranking_loss_with_cooccurrence"](r[1], y, p, scale_=1.0, margin_=1)
        record("_cooc", model, r, loss)
    np.savez_compressed(os.path.join(GOLD, f"train_ext_{tag}.npz"), **out)


def golden_losses_ext():
    """ranking_loss_with_cooccurrence (U:95-110) with the real freq_stats prior, and the two KL terms of T:809-813."""
    L = RX.loss_functions()
    F = torch.nn.functional
    kl = torch.nn.KLDivLoss(reduction="batchmean")
    g = torch.Generator().manual_seed(78)
    x = torch.randn((16, 80), generator=g) * 2.0
    xm = x + torch.randn((16, 80), generator=g) * 0.3
    y = (torch.rand((16, 80), generator=g) < 0.06).float()
    y[0, 3] = 1.0
    y[5] = 0.0                                   # a row without positives contributes nothing
    p = _cooc_p()
    out = {"x": x.numpy(), "xm": xm.numpy(), "y": y.numpy(), "cooc_p": p.numpy()}
    for name, fn in (("cooc_s1", lambda a: L["ranking_loss_with_cooccurrence"](a, y, p, scale_=1.0, margin_=1)),
                     ("cooc_s2", lambda a: L["ranking_loss_with_cooccurrence"](a, y, p)),
                     ("kl", lambda a: kl(F.log_softmax(a, dim=-1), F.softmax(xm, dim=-1))),
                     ("kl_x10000", lambda a: kl(F.log_softmax(a, dim=-1), F.softmax(xm, dim=-1)) * 10000)):
        a = x.clone().requires_grad_(True)
        b = a * 1.0
        loss = fn(b)
        loss.backward()
        out["loss_" + name] = np.float64(loss.item())
        out["grad_" + name] = a.grad.numpy()
        print(f"[losses_ext] {name}: {loss.item():.6f}", flush=True)
    np.savez_compressed(os.path.join(GOLD, "losses_ext.npz"), **out)


def build_adapter_clip(arch, sd, classnames, n_ctx, seed):
    """The reference `AdapterDenseCLIP` (TA:320-457) with the synthetic CLIP weights, seeded contexts and a seeded adapter
    (its own init is nn.Linear's default from the process RNG: replaced so the consuming side can regenerate it)."""
    ns = RX.adapter_trainer_classes()
    clip_model = RX.build_reference_clip(arch, sd)
    cfg = RX.make_cfg(arch.image_resolution, n_ctx=n_ctx)
    model = ns["AdapterDenseCLIP"](cfg, classnames, clip_model)
    w = arch.transformer_width
    with torch.no_grad():
        model.prompt_learner.ctx.copy_(synth.prompt_ctx(n_ctx, w, seed, "pos"))
        model.prompt_learner.ctx_double.copy_(synth.prompt_ctx(n_ctx, w, seed, "neg"))
        wd, wu = synth.adapter_weights(seed)
        model.adapter_text_encoder.text_adapter.fc[0].weight.copy_(wd)
        model.adapter_text_encoder.text_adapter.fc[2].weight.copy_(wu)
    for name, p in model.named_parameters():
        if "prompt_learner" not in name:
            p.requires_grad_(False)          # TA:534-536
    return model


def golden_adapter(tag, arch, classnames, n_ctx, batch_img, batch_cap, seed):
    """`AdapterDenseCLIP` (trainers/Caption_distill_double_adapter.py): test 4-tuple, train 4-tuple, ranking loss, prompt gradients."""
    sd = synth.clip_state_dict(arch, seed=0)
    L = RX.loss_functions()
    model = build_adapter_clip(arch, sd, classnames, n_ctx, seed)
    img = synth.images(batch_img, arch.image_resolution, seed)
    caps = synth.captions(batch_cap, seed, vocab=arch.vocab_size)
    y = synth.labels(batch_cap, len(classnames), seed)
    out = {"image_checksum": checksum(img), "caption_checksum": checksum(caps.float()), "label_checksum": checksum(y),
           "state_checksum": state_checksum(sd), "seed": np.int64(seed), "n_ctx": np.int64(n_ctx), "batch_img": np.int64(batch_img),
           "batch_cap": np.int64(batch_cap), "state_keys": np.array(sorted(model.state_dict().keys()))}
    with torch.no_grad():
        r = model(img, if_test=True)
    for name, t in zip(("logits", "logits_local", "neg_map", "pos_map"), r):
        out["test_" + name] = t.float().numpy()
    r = model(None, caps)
    assert len(r) == 4
    loss = L["ranking_loss"](r[0], y, scale_=1.0, margin_=1) + L["ranking_loss"](r[1], y, scale_=1.0, margin_=1)
    loss.backward()
    out["train_logits"] = r[0].detach().numpy()
    out["train_logits_local"] = r[1].detach().numpy()
    out["train_seq_feats"] = r[2].detach().numpy()[:, :2]
    out["train_text_features"] = r[3].detach().numpy()
    out["loss"] = np.float64(loss.item())
    out["grad_ctx"] = model.prompt_learner.ctx.grad.numpy()
    out["grad_ctx_double"] = model.prompt_learner.ctx_double.grad.numpy()
    print(f"[adapter_{tag}] test logits absmax {np.abs(out['test_logits']).max():.4f}; train loss {loss.item():.6f}", flush=True)
    np.savez_compressed(os.path.join(GOLD, f"adapter_{tag}.npz"), **out)


RESAMPLE_CONFIGS = {
    # the loss the trainer builds for LOSSFUNC 'dbl' (T:823-830) and the commented alternative next to it (T:831-838)
    "shipped": dict(use_sigmoid=True, reweight_func="rebalance", focal=dict(focal=False, balance_param=2.0, gamma=2), logit_reg=dict(),
                    map_param=dict(alpha=0.1, beta=10.0, gamma=0.2), loss_weight=1.0),
    "focal_reg": dict(use_sigmoid=True, reweight_func="rebalance", focal=dict(focal=True, balance_param=2.0, gamma=2),
                      logit_reg=dict(neg_scale=2.0, init_bias=0.05), map_param=dict(alpha=0.1, beta=10.0, gamma=0.2), loss_weight=1.0),
    "plain": dict(use_sigmoid=True, reweight_func=None, focal=dict(focal=False, balance_param=2.0, gamma=2), logit_reg=dict(),
                  map_param=dict(alpha=0.1, beta=10.0, gamma=0.2), loss_weight=0.5),
}


def golden_resample():
    """`ResampleLoss` (trainers/dbl.py:263-445) run on synthetic logits / labels / class frequencies."""
    import pickle
    import tempfile
    cls = RX.resample_loss_class()
    g = torch.Generator().manual_seed(79)
    x = torch.randn((16, 80), generator=g) * 2.0
    y = (torch.rand((16, 80), generator=g) < 0.06).float()
    y[0, 3] = 1.0
    y[5] = 0.0                                    # a row without positives: repeat rate 0 -> weight 1 + alpha
    class_freq = torch.randint(40, 4000, (80,), generator=g).float().numpy()
    neg_class_freq = (20000.0 - class_freq).astype(np.float32)
    out = {"x": x.numpy(), "y": y.numpy(), "class_freq": class_freq, "neg_class_freq": neg_class_freq}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "class_freq.pkl")
        with open(path, "wb") as f:
            pickle.dump({"class_freq": class_freq, "neg_class_freq": neg_class_freq}, f)
        real_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self          # dbl.py:326-341 moves its tables to the GPU; there is none here
        try:
            for name, kw in RESAMPLE_CONFIGS.items():
                fn = cls(freq_file=path, **kw)
                a = x.clone().requires_grad_(True)
                loss = fn(a * 1.0, y)                           # non-leaf copy: dbl.py:405 adds init_bias in place
                loss.backward()
                out["loss_" + name] = np.float64(loss.item())
                out["grad_" + name] = a.grad.numpy()
                print(f"[resample] {name}: {loss.item():.6f}", flush=True)
        finally:
            torch.Tensor.cuda = real_cuda
    np.savez_compressed(os.path.join(GOLD, "resample_loss.npz"), **out)


def golden_prompt_learner(classnames):
    """PromptLearner.forward(neg_prompt_wcls=False) (T:199-242: the negative / evidence prompts built WITHOUT the class-name
    tokens) and the CSC variant, on a small text width: outputs are pure concatenations, compared exactly."""
    arch = synth.tiny_rn()
    sd = synth.clip_state_dict(arch, seed=0)
    ns = RX.trainer_classes(None, arch.embed_dim)
    out = {}
    for csc in (False, True):
        clip_model = RX.build_reference_clip(arch, sd)
        cfg = RX.make_cfg(arch.image_resolution, n_ctx=4, csc=csc)
        pl = ns["PromptLearner"](cfg, TINY_CLASSES, clip_model)
        n_cls = len(TINY_CLASSES) if csc else 0
        with torch.no_grad():
            pl.ctx.copy_(synth.prompt_ctx(4, arch.transformer_width, 7, "pos", n_cls))
            pl.ctx_double.copy_(synth.prompt_ctx(4, arch.transformer_width, 7, "neg", n_cls))
            pl.ctx_evidence.copy_(synth.prompt_ctx(4, arch.transformer_width, 7, "evi"))
        sfx = "_csc" if csc else ""
        for wcls in (True, False):
            r = pl(neg_prompt_wcls=wcls)
            for nm, t in zip(("prompts", "prompts_neg", "prompts_evidence"), r[:3]):
                out[f"{nm}_wcls{int(wcls)}{sfx}"] = t.detach().numpy()
        out["name_lens" + sfx] = np.asarray(pl.name_lens, dtype=np.int64)
        out["tokenized_prompts" + sfx] = pl.tokenized_prompts.numpy()
        for k, v in pl.state_dict().items():
            out[f"state_{k}{sfx}"] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, "prompt_learner_tiny.npz"), **out)
    print("[prompt_learner] written", flush=True)


def golden_vit(tag, arch, batch, seed):
    """Reference `VisionTransformer.forward` (M:259-276) on the synthetic weights: pins the GLOBAL (class-token)
    feature of the ViT path.  The dense patch features have no reference semantics."""
    sd = synth.clip_state_dict(arch, seed=0)
    img = synth.images(batch, arch.image_resolution, seed)
    clip_model = RX.build_reference_clip(arch, sd)
    with torch.no_grad():
        g = clip_model.visual(img)
    np.savez_compressed(os.path.join(GOLD, f"vit_{tag}.npz"), global_feat=g.float().numpy(),
                        image_checksum=checksum(img), state_checksum=state_checksum(sd), batch=np.int64(batch),
                        seed=np.int64(seed))
    print(f"[vit_{tag}] global absmax={g.abs().max():.4f}", flush=True)


def golden_fusion():
    """gen_final_ans.py `fuse` / `fuse6` and the trainer's `adjust_predictions` run on synthetic window scores."""
    g = torch.Generator().manual_seed(91)
    data = torch.rand((24, 116, 80), generator=g) * 0.6 - 0.1          # window scores around the 0.2 / 0.3 thresholds
    sims = torch.rand((24, 116, 5), generator=g) * 0.4
    output = torch.rand((24, 80), generator=g)
    adj = torch.rand((80, 80), generator=g) * 100
    nums = torch.rand((80,), generator=g) * 1000 + 50
    ns = RX.fusion_functions(sims)
    p = adj / nums[:, None]
    p = p / p.sum(-1)[:, None]
    np.savez_compressed(os.path.join(GOLD, "fusion.npz"), data=data.numpy(), sims=sims.numpy(), output=output.numpy(),
                        adj=adj.numpy(), nums=nums.numpy(), fuse=ns["fuse"](data).numpy(), fuse6=ns["fuse6"](data).numpy(),
                        fuse_t05=ns["fuse"](data, threshold=0.5).numpy(),
                        adjusted=ns["adjust_predictions"](output, p, 0.5).numpy())
    print("[fusion] fuse/fuse6/adjust_predictions written", flush=True)


def golden_losses():
    L = RX.loss_functions()
    g = torch.Generator().manual_seed(77)
    x = torch.randn((16, 80), generator=g) * 2.0
    y = (torch.rand((16, 80), generator=g) < 0.06).float()
    y[0, 3] = 1.0
    y_partial = y.clone()
    y_partial[torch.rand((16, 80), generator=g) < 0.3] = 0.0
    y_partial[(y == 0) & (y_partial == 0) & (torch.rand((16, 80), generator=g) < 0.5)] = -1.0
    out = {"x": x.numpy(), "y": y.numpy(), "y_partial": y_partial.numpy()}
    for name, fn in (("ranking_s1", lambda a: L["ranking_loss"](a, y, scale_=1.0, margin_=1)),
                     ("ranking_s2", lambda a: L["ranking_loss"](a, y)),
                     ("asl", lambda a: L["ASL_loss"](a, y)),
                     ("dualcoop", lambda a: L["dualcoop_loss"](a, None, y_partial))):
        a = x.clone().requires_grad_(True)
        b = a * 1.0          # ranking_loss scales its argument in place (U:86): give it a non-leaf copy
        loss = fn(b)
        loss.backward()
        out["loss_" + name] = np.float64(loss.item())
        out["grad_" + name] = a.grad.numpy()
        print(f"[losses] {name}: {loss.item():.6f}", flush=True)
    np.savez_compressed(os.path.join(GOLD, "losses.npz"), **out)


def golden_map():
    ref_map = RX.mAP_function()
    g = torch.Generator().manual_seed(5)
    scores = torch.randn((64, 80), generator=g).numpy()
    targs = (torch.rand((64, 80), generator=g) < 0.1).float().numpy()
    targs[0, :] = 1.0          # every class has at least one positive
    np.savez_compressed(os.path.join(GOLD, "map.npz"), scores=scores, targets=targs,
                        mAP=np.float64(ref_map(targs, scores)))
    print("[map]", ref_map(targs, scores))


def main(which=None):
    assert RX.available(), "reference tree not found"
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    classnames = RX.coco_classnames()
    assert len(classnames) == 80
    steps = {
        "tokens": lambda: golden_tokens(classnames),
        "losses": golden_losses,
        "map": golden_map,
        "head_tiny": lambda: golden_head("tiny", synth.tiny_rn(), 3, TINY_CLASSES, 4, 64, 1240),
        "train_tiny": lambda: golden_train("tiny", synth.tiny_rn(), 6, TINY_CLASSES, 4, 1241),
        "head_small": lambda: golden_head("small", synth.small_rn(), 4, classnames, 16, 512, 1242),
        "head_rn50": lambda: golden_head("rn50_224", synth.RN50(224), 8, classnames, 16, 5000, 1235),
        "head_rn101": lambda: golden_head("rn101_448", synth.RN101(448), 2, classnames, 16, 2000, 1236, (True,)),
        "train_rn50": lambda: golden_train("rn50", synth.RN50(224), 4, classnames, 16, 1238),
        "fusion": golden_fusion,
        "train_ext_tiny": lambda: golden_train_ext("tiny", synth.tiny_rn(), 6, TINY_CLASSES, 4, 1243),
        "train_ext_rn50": lambda: golden_train_ext("rn50", synth.RN50(224), 4, classnames, 16, 1239),
        "losses_ext": golden_losses_ext,
        "resample": golden_resample,
        "adapter_rn50": lambda: golden_adapter("rn50", synth.RN50(224), classnames, 16, 4, 4, 1244),
        "prompt_learner": lambda: golden_prompt_learner(classnames),
        "vit_tiny": lambda: golden_vit("tiny", synth.tiny_vit(), 3, 1250),
        "vit_b16_224": lambda: golden_vit("b16_224", synth.VITB16(224), 2, 1251),
        "vit_l14_224": lambda: golden_vit("l14_224", synth.VITL14(224), 1, 1252),
    }
    for name, fn in steps.items():
        if which and name not in which:
            continue
        t0 = time.time()
        fn()
        print(f"== {name} done in {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or None)
