"""The drop-in against the REAL host (build container only: needs /root/reference): the wiring of
`Caption_distill_double.build_model` (T:737-787) re-enacted with the reference's own pieces — the reference `CLIP` module
(clip/model.py), its tokenizer, `copy.deepcopy(clip_model)`, the freeze loop (T:763-765), the optimiser over
`model.prompt_learner` only (T:773), `register_model` taking the prompt learner (T:775) — around `DenseCLIPB200` in place
of `DenseCLIP` (the `cfg.TRAIN.MODEL` switch, T:755-760).  What can be checked without a GPU: identical parameter /
buffer inventory, identical trainable set, identical attribute surface read by the trainer, checkpoint keys, and that
the forward fails loudly on CPU tensors (no fallback).  The GPU side (forward / backward under DistributedDataParallel)
is tests/test_round2_gpu.py::test_ddp_wrapped_step_equals_unwrapped."""
import copy

import pytest
import torch

from oracle import ref_extract as RX
from oracle import synth

pytestmark = pytest.mark.skipif(not RX.available(), reason="/root/reference not present")


def _pair(csc=False, ev=True):
    from oracle.make_golden import TINY_CLASSES
    from lecb200.dense_clip import DenseCLIPB200
    arch = synth.tiny_rn()
    sd = synth.clip_state_dict(arch, 0)
    cfg = RX.make_cfg(arch.image_resolution, n_ctx=4, csc=csc, use_evidence=ev)
    clip_model = RX.build_reference_clip(arch, sd)                     # load_clip_to_cpu + .float() (T:742-748)
    ns = RX.trainer_classes(synth.caption_bank(16, arch.embed_dim, 0), arch.embed_dim)
    ref = ns["DenseCLIP"](cfg, TINY_CLASSES, copy.deepcopy(clip_model), nctx=4)
    tokenize = RX.clip_package().clip.tokenize
    new = DenseCLIPB200(cfg, TINY_CLASSES, copy.deepcopy(clip_model), nctx=4, tokenizer=lambda s: tokenize(s, truncate=True))
    return ref, new


@pytest.mark.parametrize("csc", [False, True])
def test_build_model_wiring_matches(csc):
    ref, new = _pair(csc=csc)
    # same parameter and buffer inventory under the same names (state_dict keys a checkpoint of either loads into the other)
    ref_sd, new_sd = ref.state_dict(), new.state_dict()
    assert set(ref_sd) == set(new_sd), set(ref_sd) ^ set(new_sd)
    for k in ref_sd:
        assert ref_sd[k].shape == new_sd[k].shape and ref_sd[k].dtype == new_sd[k].dtype, k
    # frozen buffers built from the same tokenizer and embedding table are identical
    for k in ("prompt_learner.token_prefix", "prompt_learner.token_suffix", "prompt_learner.token_suffix_nocls"):
        assert torch.equal(ref_sd[k], new_sd[k]), k
    assert torch.equal(ref.tokenized_prompts, new.tokenized_prompts)
    assert list(ref.prompt_learner.name_lens) == list(new.prompt_learner.name_lens)
    # T:763-765: everything outside "prompt_learner" frozen -> the same trainable set (the twin is frozen by copy_params)
    for m in (ref, new):
        for name, p in m.named_parameters():
            if "prompt_learner" not in name:
                p.requires_grad_(False)
    train_ref = sorted(n for n, p in ref.named_parameters() if p.requires_grad)
    train_new = sorted(n for n, p in new.named_parameters() if p.requires_grad)
    assert train_ref == train_new and all(n.startswith("prompt_learner.") for n in train_new)
    # T:773 the optimiser sees the prompt learner only; T:775 register_model(name, model.prompt_learner, ...)
    opt_ref = torch.optim.SGD(ref.prompt_learner.parameters(), lr=0.002)
    opt_new = torch.optim.SGD(new.prompt_learner.parameters(), lr=0.002)
    assert [tuple(p.shape) for p in opt_ref.param_groups[0]["params"]] == [tuple(p.shape) for p in opt_new.param_groups[0]["params"]]
    assert set(ref.prompt_learner.state_dict()) == set(new.prompt_learner.state_dict())
    # attributes the trainer / test loop read (T:568, T:775, T:906-938)
    for attr in ("prompt_learner", "prompt_learner_m", "tokenized_prompts", "text_encoder", "model", "logit_scale", "dtype",
                 "cfg", "prompt_text_features", "model_pairs", "copy_params", "_momentum_update", "encode_image",
                 "v_linear_weight", "v_linear_bias", "c_linear_weight", "c_linear_bias"):
        assert hasattr(new, attr) == hasattr(ref, attr) is True, attr
    assert new.v_linear_weight is new.model.visual.attnpool.v_proj.weight              # aliases, T:370-373


def test_prompt_learner_forward_is_the_reference_expression():
    """Same parameters in, same prompt embeddings out (pure torch in both): what the text tower kernels then consume."""
    ref, new = _pair()
    new.prompt_learner.load_state_dict(ref.prompt_learner.state_dict())
    for wcls in (True, False):
        a, b = ref.prompt_learner(neg_prompt_wcls=wcls), new.prompt_learner(neg_prompt_wcls=wcls)
        assert len(a) == len(b) == 6
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_checkpoint_of_the_reference_loads_into_the_drop_in(tmp_path):
    """save_model / load_model interchange on the reference's own torchtools functions (T:906-938)."""
    from lecb200 import checkpoint as CK
    ref, new = _pair()
    tools = RX.checkpoint_functions()
    with torch.no_grad():
        for p in ref.prompt_learner.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    tools["save_checkpoint"]({"state_dict": ref.prompt_learner.state_dict(), "epoch": 3, "optimizer": None, "scheduler": None},
                             str(tmp_path / "prompt_learner"), model_name="model.pth.tar-3")
    ckpt = CK.load_checkpoint(str(tmp_path / "prompt_learner" / "model.pth.tar-3"))
    sd = {k: v for k, v in ckpt["state_dict"].items() if "token_prefix" not in k and "token_suffix" not in k}      # T:927-935
    new.prompt_learner.load_state_dict(sd, strict=False)
    for (n1, p1), (n2, p2) in zip(ref.prompt_learner.named_parameters(), new.prompt_learner.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)


def test_forward_on_cpu_fails_loudly():
    from lecb200 import LecbError
    _, new = _pair()
    with pytest.raises(LecbError):
        new(torch.zeros((1, 3, 64, 64)), if_test=True)
    with pytest.raises(LecbError):
        new(None, torch.zeros((2, 77), dtype=torch.long))
