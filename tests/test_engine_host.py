"""Host-side weight preparation of the visual engine (no GPU): the projection-shortcut tail of a stage's first Bottleneck
(M:46-52: bn3(conv3(y)) + downsample(x), ReLU) is handed to lecb_gemm_bf16_dual as ONE weight [W3 | Wd] and one bias b3 + bd —
checked here against the two folded convolutions evaluated separately in fp32 torch."""
import torch

import lecb200  # noqa: F401
from lecb200 import engine, synth


def _engine(monkeypatch=None, no_dual=False):
    arch = synth.RN50(224)
    sd = synth.clip_state_dict(arch, 0)
    if no_dual:
        monkeypatch.setenv("LECB_NO_DUAL", "1")
    return engine.VisualRN(sd, arch.vision_layers, arch.vision_width, arch.vision_width * 32 // 64, arch.embed_dim, "cpu")


def test_shortcut_tail_weights_are_the_concatenation():
    eng = _engine()
    entries = [blk for blk in eng.blocks if "ds" in blk]
    assert len(entries) == 4                                           # one projection shortcut per stage
    for blk in entries:
        assert "c3ds" in blk, "64-channel granularity holds for every CLIP ResNet width"
        (w3, b3), (wd, bd), (w, b) = blk["c3"], blk["ds"], blk["c3ds"]
        k1, k2 = w3.shape[1], wd.shape[1]
        assert w.dtype == torch.bfloat16 and w.is_contiguous() and tuple(w.shape) == (w3.shape[0], k1 + k2)
        assert torch.equal(w[:, :k1], w3) and torch.equal(w[:, k1:], wd)
        g = torch.Generator().manual_seed(k1)
        y = torch.randn((64, k1), generator=g).bfloat16().float()
        x = torch.randn((64, k2), generator=g).bfloat16().float()
        two = (y @ w3.float().t() + b3) + (x @ wd.float().t() + bd)    # the reference's two convolutions + add
        one = torch.cat([y, x], 1) @ w.float().t() + b                 # what the dual-operand GEMM computes
        torch.testing.assert_close(one, two, rtol=1e-5, atol=1e-5)
    # identity blocks keep the residual form
    assert all("c3ds" not in blk for blk in eng.blocks if "ds" not in blk)


def test_shortcut_tail_switch(monkeypatch):
    eng = _engine(monkeypatch, no_dual=True)
    assert all("c3ds" not in blk for blk in eng.blocks)


def test_narrow_towers_keep_the_two_gemm_form():
    arch = synth.tiny_rn()                       # width 8: bottleneck widths 8 .. 64, shortcut inputs 8 .. 128
    sd = synth.clip_state_dict(arch, 0)
    eng = engine.VisualRN(sd, arch.vision_layers, arch.vision_width, max(1, arch.vision_width * 32 // 64), arch.embed_dim, "cpu")
    entries = [blk for blk in eng.blocks if "ds" in blk]
    assert len(entries) == 4
    for blk in entries:
        aligned = blk["c3"][0].shape[1] % 64 == 0 and blk["ds"][0].shape[1] % 64 == 0
        assert ("c3ds" in blk) == aligned
    assert not all("c3ds" in blk for blk in entries)
