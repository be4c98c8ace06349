"""Prompt checkpoints and test-time dumps in the reference's on-disk formats, on the CPU.

Format checks need nothing but torch; the cross-checks run the reference's own `save_checkpoint` / `load_checkpoint` /
`load_pretrained_weights` (AST-extracted from dassl/utils/torchtools.py) against lecb200.checkpoint in both directions and
only run where /root/reference exists."""
import importlib.util
import os

import pytest
import torch
import torch.nn as nn

from oracle import ref_extract as RX

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ckpt_module():
    # loaded by path: importing the package itself needs the CUDA library, these formats do not
    spec = importlib.util.spec_from_file_location(
        "_lecb200_checkpoint", os.path.join(ROOT, "language-enhanced-clip-for-multi-label-image-recognition_b200", "checkpoint.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CK = _ckpt_module()


class _Prompts(nn.Module):
    """Parameter / buffer names of the reference PromptLearner (T:104-197)."""

    def __init__(self, n_ctx=4, width=8, n_cls=3, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        for n in ("ctx", "ctx_double", "ctx_evidence"):
            setattr(self, n, nn.Parameter(torch.randn((n_ctx, width), generator=g) * 0.02))
        for n, v in (("temperature", 3.0), ("spatial_T", 3.0), ("ranking_scale", 4.0)):
            setattr(self, n, nn.Parameter(torch.tensor(v)))
        self.register_buffer("token_prefix", torch.randn((n_cls, 1, width), generator=g))
        self.register_buffer("token_suffix", torch.randn((n_cls, 5, width), generator=g))
        self.register_buffer("token_suffix_nocls", torch.randn((n_cls, 5, width), generator=g))
        self.resets = 0

    def reset_prompt_cache(self):
        self.resets += 1


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_save_model_layout_and_roundtrip(tmp_path):
    m = _Prompts(seed=1)
    opt = torch.optim.SGD(m.parameters(), lr=0.002, momentum=0.9)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, 50)
    m.ctx.sum().backward()
    opt.step()
    sched.step()
    paths = CK.save_model({"double": (m, opt, sched)}, epoch=9, directory=str(tmp_path), is_best=True)
    d = tmp_path / "double"
    assert paths["double"] == str(d / "model.pth.tar-10")                   # the reference stores epoch + 1
    assert (d / "checkpoint").read_text() == "model.pth.tar-10\n"
    assert (d / "model-best.pth.tar").exists()
    raw = torch.load(d / "model.pth.tar-10", weights_only=False)
    assert set(raw) == {"state_dict", "epoch", "optimizer", "scheduler"} and raw["epoch"] == 10
    assert raw["optimizer"]["param_groups"][0]["momentum"] == 0.9 and raw["scheduler"]["T_max"] == 50

    m2 = _Prompts(seed=2)
    before = {k: v.clone() for k, v in m2.state_dict().items()}
    epochs = CK.load_model({"double": m2}, str(tmp_path), epoch=10)
    assert epochs == {"double": 10} and m2.resets == 1
    for k, v in m2.state_dict().items():
        if k in ("token_prefix", "token_suffix"):                            # fixed token vectors are ignored (T:928-932)
            assert torch.equal(v, before[k])
        else:
            assert torch.equal(v, m.state_dict()[k]), k
    # default = the best model file name without an epoch suffix
    os.replace(d / "model-best.pth.tar", d / "model.pth.tar")
    assert CK.load_model({"double": _Prompts(seed=3)}, str(tmp_path)) == {"double": 10}


def test_load_errors_match_reference_behaviour(tmp_path):
    assert CK.load_model({"double": _Prompts()}, "") is None                 # skipped, T:907-909
    with pytest.raises(FileNotFoundError, match="Model not found at"):
        CK.load_model({"double": _Prompts()}, str(tmp_path), epoch=3)
    with pytest.raises(ValueError):
        CK.load_checkpoint(None)
    with pytest.raises(FileNotFoundError, match="File is not found at"):
        CK.load_checkpoint(str(tmp_path / "nope.pth.tar"))


def test_module_prefix_and_pretrained_weights(tmp_path):
    m = _Prompts(seed=4)
    state = {"state_dict": {"module." + k: v for k, v in m.state_dict().items()}, "epoch": 1, "optimizer": None, "scheduler": None}
    f = CK.save_checkpoint(state, str(tmp_path / "p"), model_name="model-best.pth.tar")
    assert all(not k.startswith("module.") for k in torch.load(f, weights_only=False)["state_dict"])
    other = _Prompts(n_ctx=6, seed=5)                                        # ctx* differ in size -> discarded, scalars load
    matched, discarded = CK.load_pretrained_weights(other, f)
    assert set(discarded) == {"ctx", "ctx_double", "ctx_evidence"} and "temperature" in matched
    assert other.resets == 1 and torch.equal(other.token_prefix, m.token_prefix)


def test_logit_dump_and_sim_matrix_formats(tmp_path):
    g = torch.Generator().manual_seed(7)
    batches = [torch.randn((4, 80), generator=g) for _ in range(3)]
    blocks = [torch.randn((4, 116, 80), generator=g) for _ in range(3)]
    p = tmp_path / "train_output" / "data.pth"
    CK.save_logit_dump(str(p), {"double": {"output": batches, "output_pos": batches, "output_blocks": blocks, "output_pos_blocks": blocks},
                                "ema": {"output": torch.cat(batches), "output_pos": torch.cat(batches)}})
    d = CK.load_logit_dump(str(p))
    assert set(d) == {"double", "ema"} and set(d["ema"]) == {"output", "output_pos"}
    assert d["double"]["output"].shape == (12, 80) and d["double"]["output_blocks"].shape == (12, 116, 80)
    assert torch.equal(d["double"]["output"], torch.cat(batches))
    with pytest.raises(KeyError):
        CK.save_logit_dump(str(tmp_path / "bad.pth"), {"double": {"output": batches}})
    s = tmp_path / "train_output" / "sim_matrix_B.pth"
    CK.save_sim_matrix(str(s), [torch.ones((4, 10))] * 3, [torch.ones((4, 116, 10))] * 3)
    CK.save_sim_matrix(str(s), torch.zeros((1, 10)), torch.zeros((1, 116, 10)))          # kept: the reference saves it once
    sm = torch.load(s, weights_only=False)
    assert set(sm) == {"sims_all", "sims_blocks_all"} and sm["sims_all"].shape == (12, 10) and sm["sims_blocks_all"].shape == (12, 116, 10)


@pytest.mark.skipif(not RX.available(), reason="/root/reference not present")
def test_files_interchange_with_the_reference_functions(tmp_path):
    REF = RX.checkpoint_functions()
    m = _Prompts(seed=11)
    opt = torch.optim.SGD(m.parameters(), lr=0.002)
    state = lambda: {"state_dict": {"module." + k: v.clone() for k, v in m.state_dict().items()}, "epoch": 7,
                     "optimizer": opt.state_dict(), "scheduler": None}
    # reference writes, we read
    REF["save_checkpoint"](state(), str(tmp_path / "ref"), is_best=True)
    ours = CK.load_checkpoint(str(tmp_path / "ref" / "model.pth.tar-7"))
    _same(ours["state_dict"], dict(m.state_dict()))
    assert ours["epoch"] == 7 and (tmp_path / "ref" / "checkpoint").read_text() == "model.pth.tar-7\n"
    # we write, the reference reads; byte-level layout of the directory is the same
    CK.save_checkpoint(state(), str(tmp_path / "ours"), is_best=True)
    assert sorted(os.listdir(tmp_path / "ours")) == sorted(os.listdir(tmp_path / "ref"))
    assert (tmp_path / "ours" / "checkpoint").read_bytes() == (tmp_path / "ref" / "checkpoint").read_bytes()
    theirs = REF["load_checkpoint"](str(tmp_path / "ours" / "model-best.pth.tar"))
    _same(theirs["state_dict"], dict(m.state_dict()))
    assert theirs["optimizer"]["param_groups"] == opt.state_dict()["param_groups"]
    # load_pretrained_weights: same layers matched / discarded, same resulting weights
    a, b = _Prompts(n_ctx=6, seed=12), _Prompts(n_ctx=6, seed=12)
    REF["load_pretrained_weights"](a, str(tmp_path / "ours" / "model.pth.tar-7"))
    CK.load_pretrained_weights(b, str(tmp_path / "ref" / "model.pth.tar-7"))
    _same(dict(a.state_dict()), dict(b.state_dict()))
    with pytest.raises(FileNotFoundError):
        REF["load_checkpoint"](str(tmp_path / "nope"))
