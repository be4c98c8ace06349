"""Helpers shared by the -m gpu parity tests."""
import numpy as np
import torch

from oracle import restatement as R
from oracle import ref_extract  # noqa: F401  (only its Cfg helper is used; never touches /root/reference here)
from oracle.ref_extract import make_cfg

from . import _cases as C

LOGIT_TOL = 1e-2        # north_star: max abs logit error <= 1e-2 (bf16 operands, fp32 accumulation)


def build_model(case, use_evidence, bank=None, tag="coco", **cfg_kw):
    """lecb200 DenseCLIPB200 on cuda:0 holding the case's synthetic weights and prompt contexts.
    cfg_kw: make_cfg switches (csc=, ema=, learn_scale=, ...)."""
    from lecb200.clip_model import CLIPParams
    from lecb200.dense_clip import DenseCLIPB200
    arch = case["arch"]
    clip = CLIPParams(*arch.ctor_args())
    missing, unexpected = clip.load_state_dict(case["sd"], strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    clip = clip.float().cuda().eval()
    toks, n_ctx, names = C.tokens_for(tag)
    cfg = make_cfg(arch.image_resolution, n_ctx=n_ctx, use_evidence=use_evidence, **cfg_kw)
    model = DenseCLIPB200(cfg, names, clip, caption_bank=bank, tokenized_prompts=toks).cuda()
    with torch.no_grad():
        model.prompt_learner.ctx.copy_(case["pl_state"]["ctx"])
        model.prompt_learner.ctx_double.copy_(case["pl_state"]["ctx_double"])
        model.prompt_learner.ctx_evidence.copy_(case["pl_state"]["ctx_evidence"])
    model.copy_params()
    return model


def rel_err(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return ((got - want).abs().max() / (want.abs().max() + 1e-12)).item()


def topk_sets_match(got, want, k, tol):
    """Top-k label sets identical, up to classes whose reference score ties the k-th score within tol."""
    got, want = np.asarray(got), np.asarray(want)
    for r in range(want.shape[0]):
        sg, sw = set(np.argsort(-got[r])[:k]), set(np.argsort(-want[r])[:k])
        kth = np.sort(want[r])[::-1][k - 1]
        for c in sg ^ sw:
            if abs(want[r, c] - kth) > tol:
                return False
    return True
