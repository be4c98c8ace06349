"""Shared builders for parity cases: regenerate seeded inputs/weights and pair them with the
reference-generated fixtures in tests/golden/ (written by oracle/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import restatement as R
from oracle import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

HEAD_CASES = {
    # tag: (arch factory, evidence modes present in the fixture)
    "tiny": (synth.tiny_rn, (False, True)),
    "small": (synth.small_rn, (False, True)),
    "rn50_224": (lambda: synth.RN50(224), (False, True)),
    "rn101_448": (lambda: synth.RN101(448), (True,)),
}
TRAIN_CASES = {"tiny": synth.tiny_rn, "rn50": lambda: synth.RN50(224)}


def load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def checksum(t):
    t = t.detach().double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64) % 97 + 1
    return np.array([t.sum().item(), (t * w).sum().item(), t.abs().max().item()], dtype=np.float64)


def tokens_for(tag):
    tk = load("prompt_tokens_coco80.npz")
    if tag == "tiny":
        return torch.from_numpy(tk["tiny_tokens"]), int(tk["tiny_n_ctx"]), [str(s) for s in tk["tiny_classnames"]]
    return torch.from_numpy(tk["tokens"]), int(tk["n_ctx"]), [str(s) for s in tk["classnames"]]


def prompt_state(sd, arch, tag, seed):
    toks, n_ctx, _ = tokens_for(tag)
    w = arch.transformer_width
    ctx = [synth.prompt_ctx(n_ctx, w, seed, t) for t in ("pos", "neg", "evi")]
    return R.prompt_learner_state(sd, toks, n_ctx, *ctx), toks, n_ctx


def head_case(tag):
    """-> dict(arch, sd, image, bank, pl_state, tokens, gold)"""
    g = load(f"head_{tag}.npz")
    arch = HEAD_CASES[tag][0]()
    seed = int(g["seed"])
    sd = synth.clip_state_dict(arch, 0)
    image = synth.images(int(g["batch"]), arch.image_resolution, seed)
    bank = synth.caption_bank(int(g["bank_rows"]), arch.embed_dim, seed)
    np.testing.assert_allclose(checksum(image), g["image_checksum"], rtol=1e-12, err_msg="RNG drift: images")
    np.testing.assert_allclose(checksum(bank.float()), g["bank_checksum"], rtol=1e-12, err_msg="RNG drift: bank")
    pl, toks, n_ctx = prompt_state(sd, arch, "tiny" if tag == "tiny" else "coco", seed)
    return dict(arch=arch, sd=sd, image=image, bank=bank, pl_state=pl, tokens=toks, n_ctx=n_ctx,
                gold=g, seed=seed, modes=HEAD_CASES[tag][1])


def train_case(tag):
    g = load(f"train_{tag}.npz")
    arch = TRAIN_CASES[tag]()
    seed = int(g["seed"])
    sd = synth.clip_state_dict(arch, 0)
    _, _, names = tokens_for("tiny" if tag == "tiny" else "coco")
    caps = synth.captions(int(g["batch"]), seed, vocab=arch.vocab_size)
    y = synth.labels(int(g["batch"]), len(names), seed)
    np.testing.assert_allclose(checksum(caps.float()), g["caption_checksum"], rtol=1e-12, err_msg="RNG drift: captions")
    np.testing.assert_allclose(checksum(y), g["label_checksum"], rtol=1e-12, err_msg="RNG drift: labels")
    pl, toks, n_ctx = prompt_state(sd, arch, "tiny" if tag == "tiny" else "coco", seed)
    return dict(arch=arch, sd=sd, captions=caps, labels=y, pl_state=pl, tokens=toks, n_ctx=n_ctx, gold=g, seed=seed)
