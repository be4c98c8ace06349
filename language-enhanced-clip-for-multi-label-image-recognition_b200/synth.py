"""Seeded synthetic CLIP weights and inputs (pure torch RNG; used by bench.py, the tests and the oracle).

There is no network, so no OpenAI checkpoint: every parity case runs on random-init weights.
The reference's own initialiser (`clip/model.py:335-362`, `CLIP.initialize_parameters`) zeroes every
`bn3.weight` (model.py:347-350) and leaves BN running stats at (0, 1), which silences all residual
branches (SURVEY §8c deviation 3).  This factory therefore builds a *state_dict* with the OpenAI CLIP
key names / shapes (what `clip/model.py:435-472` `build_model` consumes) where each tensor is drawn
from its own generator seeded by crc32(key) ^ seed.  The same dict is loaded into the reference
`CLIP` module (oracle/make_golden.py, in the build container) and into the B200 engine, so both sides
see bit-identical fp32 weights without shipping 400 MB of fixtures.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from dataclasses import dataclass
from typing import Tuple, Union

import torch


@dataclass(frozen=True)
class ClipArch:
    """Constructor arguments of the reference `CLIP` (clip/model.py:280-293)."""
    embed_dim: int
    image_resolution: int
    vision_layers: Union[Tuple[int, int, int, int], int]
    vision_width: int
    vision_patch_size: Union[int, None]
    context_length: int = 77
    vocab_size: int = 49408
    transformer_width: int = 512
    transformer_heads: int = 8
    transformer_layers: int = 12

    def ctor_args(self):
        return (self.embed_dim, self.image_resolution, self.vision_layers, self.vision_width,
                self.vision_patch_size, self.context_length, self.vocab_size,
                self.transformer_width, self.transformer_heads, self.transformer_layers)

    @property
    def is_resnet(self):
        return isinstance(self.vision_layers, (tuple, list))


def RN50(res=224):
    return ClipArch(1024, res, (3, 4, 6, 3), 64, None)


def RN101(res=448):
    return ClipArch(512, res, (3, 4, 23, 3), 64, None)


def VITB16(res=448):
    return ClipArch(512, res, 12, 768, 16)


def VITL14(res=448):
    return ClipArch(768, res, 24, 1024, 14, transformer_width=768, transformer_heads=12)


def tiny_vit(res=96, patch=16, width=128, layers=3, embed_dim=256, text_width=64, text_layers=2):
    """A small VisionTransformer-shaped arch (head dim 64 like every CLIP ViT) for fast tests."""
    return ClipArch(embed_dim, res, layers, width, patch, 77, 49408, text_width, max(1, text_width // 64), text_layers)


def tiny_rn(res=64, embed_dim=64, width=8, layers=(1, 1, 1, 1), text_width=64, text_layers=2, vocab=49408):
    """A small ModifiedResNet-shaped arch for fast CPU tests (same code path, toy sizes)."""
    return ClipArch(embed_dim, res, tuple(layers), width, None, 77, vocab, text_width,
                    max(1, text_width // 64), text_layers)


def small_rn(res=128, embed_dim=512, layers=(1, 1, 1, 1)):
    """Full-width (64) ModifiedResNet with one block per stage: every channel count the B200 kernels
    see in RN50/RN101 (32..2048) at a fraction of the depth — the GPU parity-test workhorse."""
    return ClipArch(embed_dim, res, tuple(layers), 64, None)


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFFFFFFFFFF)
    return g


def _normal(key, seed, shape, std=1.0, mean=0.0):
    return torch.randn(shape, generator=_gen(key, seed), dtype=torch.float32) * std + mean


def _uniform(key, seed, shape, lo, hi):
    return torch.rand(shape, generator=_gen(key, seed), dtype=torch.float32) * (hi - lo) + lo


def _bn(sd, prefix, c, seed, gamma=(0.5, 1.5)):
    sd[prefix + ".weight"] = _uniform(prefix + ".weight", seed, (c,), *gamma)
    sd[prefix + ".bias"] = _normal(prefix + ".bias", seed, (c,), 0.1)
    sd[prefix + ".running_mean"] = _normal(prefix + ".running_mean", seed, (c,), 0.1)
    sd[prefix + ".running_var"] = _uniform(prefix + ".running_var", seed, (c,), 0.5, 1.5)
    sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _conv(sd, key, cout, cin, k, seed):
    fan_in = cin * k * k
    sd[key] = _normal(key, seed, (cout, cin, k, k), (2.0 / fan_in) ** 0.5)


def _linear(sd, prefix, cout, cin, seed, std=None, bias_std=0.02):
    sd[prefix + ".weight"] = _normal(prefix + ".weight", seed, (cout, cin), std if std else cin ** -0.5)
    sd[prefix + ".bias"] = _normal(prefix + ".bias", seed, (cout,), bias_std)


def _ln(sd, prefix, c, seed):
    sd[prefix + ".weight"] = _normal(prefix + ".weight", seed, (c,), 0.1, 1.0)
    sd[prefix + ".bias"] = _normal(prefix + ".bias", seed, (c,), 0.05)


def _transformer(sd, prefix, width, layers, seed):
    proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
    for i in range(layers):
        p = f"{prefix}.resblocks.{i}"
        sd[p + ".attn.in_proj_weight"] = _normal(p + ".attn.in_proj_weight", seed, (3 * width, width), width ** -0.5)
        sd[p + ".attn.in_proj_bias"] = _normal(p + ".attn.in_proj_bias", seed, (3 * width,), 0.02)
        _linear(sd, p + ".attn.out_proj", width, width, seed, proj_std)
        _ln(sd, p + ".ln_1", width, seed)
        _linear(sd, p + ".mlp.c_fc", 4 * width, width, seed, (2 * width) ** -0.5)
        _linear(sd, p + ".mlp.c_proj", width, 4 * width, seed, proj_std)
        _ln(sd, p + ".ln_2", width, seed)


def clip_state_dict(arch: ClipArch, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """fp32 state_dict with the key set of the reference `CLIP` module (clip/model.py:279-333)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    if arch.is_resnet:
        w = arch.vision_width
        _conv(sd, "visual.conv1.weight", w // 2, 3, 3, seed)
        _bn(sd, "visual.bn1", w // 2, seed)
        _conv(sd, "visual.conv2.weight", w // 2, w // 2, 3, seed)
        _bn(sd, "visual.bn2", w // 2, seed)
        _conv(sd, "visual.conv3.weight", w, w // 2, 3, seed)
        _bn(sd, "visual.bn3", w, seed)
        inplanes = w
        for li, (planes, blocks) in enumerate(zip((w, 2 * w, 4 * w, 8 * w), arch.vision_layers), start=1):
            for bi in range(blocks):
                p = f"visual.layer{li}.{bi}"
                stride = 2 if (li > 1 and bi == 0) else 1
                _conv(sd, p + ".conv1.weight", planes, inplanes, 1, seed)
                _bn(sd, p + ".bn1", planes, seed)
                _conv(sd, p + ".conv2.weight", planes, planes, 3, seed)
                _bn(sd, p + ".bn2", planes, seed)
                _conv(sd, p + ".conv3.weight", planes * 4, planes, 1, seed)
                # keep the residual stream from exploding over 33 blocks: small (but live) branch gain
                _bn(sd, p + ".bn3", planes * 4, seed, gamma=(0.1, 0.3))
                if stride > 1 or inplanes != planes * 4:
                    _conv(sd, p + ".downsample.0.weight", planes * 4, inplanes, 1, seed)
                    _bn(sd, p + ".downsample.1", planes * 4, seed)
                inplanes = planes * 4
        cv = w * 32
        sp = arch.image_resolution // 32
        sd["visual.attnpool.positional_embedding"] = _normal("visual.attnpool.positional_embedding", seed,
                                                             (sp * sp + 1, cv), cv ** -0.5)
        # q/k at half the reference init std (M:341-345) so that random-weight attention scores stay O(1-5)
        # instead of saturating the softmax (an ill-conditioned argmax no real checkpoint exhibits)
        for nm, gain in (("k_proj", 0.5), ("q_proj", 0.5), ("v_proj", 1.0)):
            _linear(sd, f"visual.attnpool.{nm}", cv, cv, seed, gain * cv ** -0.5, 0.1)
        _linear(sd, "visual.attnpool.c_proj", arch.embed_dim, cv, seed, cv ** -0.5, 0.1)
    else:
        w = arch.vision_width
        ps = arch.vision_patch_size
        grid = arch.image_resolution // ps
        sd["visual.class_embedding"] = _normal("visual.class_embedding", seed, (w,), w ** -0.5)
        sd["visual.positional_embedding"] = _normal("visual.positional_embedding", seed, (grid * grid + 1, w), w ** -0.5)
        sd["visual.proj"] = _normal("visual.proj", seed, (w, arch.embed_dim), w ** -0.5)
        sd["visual.conv1.weight"] = _normal("visual.conv1.weight", seed, (w, 3, ps, ps), (3 * ps * ps) ** -0.5)
        _ln(sd, "visual.ln_pre", w, seed)
        _transformer(sd, "visual.transformer", w, arch.vision_layers, seed)
        _ln(sd, "visual.ln_post", w, seed)
    tw = arch.transformer_width
    _transformer(sd, "transformer", tw, arch.transformer_layers, seed)
    sd["token_embedding.weight"] = _normal("token_embedding.weight", seed, (arch.vocab_size, tw), 0.02)
    sd["positional_embedding"] = _normal("positional_embedding", seed, (arch.context_length, tw), 0.01)
    _ln(sd, "ln_final", tw, seed)
    sd["text_projection"] = _normal("text_projection", seed, (tw, arch.embed_dim), tw ** -0.5)
    sd["logit_scale"] = torch.tensor(2.6592600, dtype=torch.float32)
    return sd


def images(batch: int, res: int, seed: int) -> torch.Tensor:
    """`image ~ N(0,1)` [B,3,res,res] fp32 (≈ CLIP-normalised pixels), SURVEY §8(d)."""
    return torch.randn((batch, 3, res, res), generator=_gen("images", seed), dtype=torch.float32)


def caption_bank(n: int, dim: int, seed: int, dtype=torch.float16) -> torch.Tensor:
    """Row-normalised random bank standing in for the 220k-caption feature file read at
    Caption_distill_double.py:35-36 (produced by generate_caption_text_features.py:72-96, fp16)."""
    x = torch.randn((n, dim), generator=_gen("bank", seed), dtype=torch.float32)
    return (x / x.norm(dim=-1, keepdim=True)).to(dtype)


def prompt_ctx(n_ctx: int, dim: int, seed: int, tag: str, n_cls: int = 0) -> torch.Tensor:
    """`ctx* ~ N(0, 0.02)` like Caption_distill_double.py:130-151 ([n_ctx,dim] or CSC [n_cls,n_ctx,dim])."""
    shape = (n_cls, n_ctx, dim) if n_cls else (n_ctx, dim)
    return _normal("ctx/" + tag, seed, shape, 0.02)


def captions(batch: int, seed: int, vocab: int = 49408, ctx_len: int = 77, dense: bool = False) -> torch.Tensor:
    """Token ids [B,77] int64: SOT, body ids in [1, vocab-3], EOT (= max id) then zeros.
    Length drawn from a clipped Gaussian matching the shipped captions (mean 20, max 37)."""
    g = _gen("captions", seed)
    sot, eot = vocab - 2, vocab - 1
    out = torch.zeros((batch, ctx_len), dtype=torch.long)
    for b in range(batch):
        if dense:
            n = ctx_len
        else:
            n = int(torch.randn((), generator=g).item() * 5.0 + 20.0)
            n = max(9, min(ctx_len, n))
        body = torch.randint(1, vocab - 2, (n - 2,), generator=g)
        out[b, 0] = sot
        out[b, 1:n - 1] = body
        out[b, n - 1] = eot
    return out


def labels(batch: int, n_cls: int, seed: int) -> torch.Tensor:
    """Multi-hot labels with 1–5 positives per row, float32 [B,K]."""
    g = _gen("labels", seed)
    y = torch.zeros((batch, n_cls), dtype=torch.float32)
    for b in range(batch):
        k = int(torch.randint(1, 6, (), generator=g).item())
        idx = torch.randperm(n_cls, generator=g)[:k]
        y[b, idx] = 1.0
    return y


def adapter_weights(seed: int, c_in: int = 512, reduction: int = 4):
    """Seeded weights of the adapter trainer's `Adapter(512, 4)` (two bias-free linears, Caption_distill_double_adapter.py:304-317):
    (down [c_in/r, c_in], up [c_in, c_in/r]), scaled so the residual branch is live but O(1)."""
    down = _normal("adapter/down", seed, (c_in // reduction, c_in), c_in ** -0.5)
    up = _normal("adapter/up", seed, (c_in, c_in // reduction), (c_in // reduction) ** -0.5)
    return down, up
