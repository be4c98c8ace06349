"""ViT visual tower (BASELINE configs 3, 5) on the B200 vs the CPU oracle and the reference-generated goldens.

Parity status: the GLOBAL (class-token) feature is the reference `VisionTransformer.forward` (M:259-276) and is
pinned by tests/golden/vit_*.npz; the DENSE patch features are this repo's definition (the reference `DenseCLIP`
cannot wrap a ViT) and are compared with oracle/restatement.py `vit_dense` only ("parity vs repo oracle").
Gate: max abs logit error <= 1e-2 like the ModifiedResNet path."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import synth

from . import _cases as C
from ._gpu_common import LOGIT_TOL, build_model, rel_err, topk_sets_match

pytestmark = pytest.mark.gpu


def _case(arch, batch, seed, tag):
    sd = synth.clip_state_dict(arch, 0)
    pl, toks, n_ctx = C.prompt_state(sd, arch, tag, seed)
    return dict(arch=arch, sd=sd, image=synth.images(batch, arch.image_resolution, seed), pl_state=pl, tokens=toks, n_ctx=n_ctx)


@pytest.mark.parametrize("tag,arch_fn", [("tiny", synth.tiny_vit), ("b16_224", lambda: synth.VITB16(224)),
                                         ("l14_224", lambda: synth.VITL14(224))])
def test_vit_tokens_match_golden_and_oracle(tag, arch_fn):
    g = C.load(f"vit_{tag}.npz")
    arch = arch_fn()
    c = _case(arch, int(g["batch"]), int(g["seed"]), "tiny" if tag == "tiny" else "coco")
    model = build_model(c, use_evidence=False, tag="tiny" if tag == "tiny" else "coco")
    eng = model.visual_engine()
    feat, ssq, t = eng.tokens(c["image"].cuda())
    torch.cuda.synchronize()
    b = c["image"].shape[0]
    feat3 = feat.float().view(b, t, -1).cpu()
    e = rel_err(feat3[:, 0], torch.from_numpy(g["global_feat"]))
    print(f"[vit {tag}] class-token feature rel err vs REFERENCE golden {e:.4f}")
    assert e < 3e-2
    with torch.no_grad():
        og, ol = R.vit_dense(c["sd"], c["image"], arch.vision_patch_size, arch.vision_width // 64)
    e = rel_err(feat3[:, 1:].permute(1, 0, 2), ol)
    print(f"[vit {tag}] dense patch features rel err vs repo oracle {e:.4f}")
    assert e < 3e-2
    want = feat.float().pow(2).sum(-1)
    assert ((ssq - want).abs() / want).max().item() < 1e-3


@pytest.mark.parametrize("tag,arch_fn,batch", [("tiny", synth.tiny_vit, 3), ("b16_448", lambda: synth.VITB16(448), 2),
                                               ("l14_448", lambda: synth.VITL14(448), 1)])
def test_vit_forward_test_matches_oracle(tag, arch_fn, batch):
    arch = arch_fn()
    ptag = "tiny" if tag == "tiny" else "coco"
    c = _case(arch, batch, 1260, ptag)
    for ev in (False, True):
        model = build_model(c, use_evidence=ev, tag=ptag)
        out = model(c["image"].cuda(), if_test=True)
        torch.cuda.synchronize()
        logits, logits_local, neg_map, pos_map, scores = [None if t is None else t.float().cpu() for t in out]
        with torch.no_grad():
            ref = R.dense_clip_test_vit(c["sd"], arch, c["image"], c["pl_state"], c["tokens"], use_evidence=ev)
        errs = [(a - r).abs().max().item() for a, r in zip((logits, logits_local, neg_map, pos_map), ref[:4])]
        print(f"[vit {tag} ev={ev}] max abs err logits_ {errs[0]:.5f} logits_local {errs[1]:.5f} neg_map {errs[2]:.5f} pos_map {errs[3]:.5f}")
        assert neg_map.shape == ref[2].shape and pos_map.shape == ref[3].shape
        assert max(errs) <= LOGIT_TOL
        assert topk_sets_match(logits.numpy(), ref[0].numpy(), 5, 2 * LOGIT_TOL)
        assert topk_sets_match(logits_local.numpy(), ref[1].numpy(), 5, 2 * LOGIT_TOL)
