"""Parameter containers with the OpenAI-CLIP checkpoint key layout (what the reference's
`clip.build_model` consumes, clip/model.py:435-472).  They only HOLD weights: nothing here computes —
the B200 engine (engine.py) reads `state_dict()` and runs its own kernels.  A reference `CLIP`
instance can be passed to `DenseCLIPB200` in their place (same keys, same attribute names:
`.visual.attnpool`, `.visual.input_resolution`, `.transformer`, `.token_embedding`,
`.positional_embedding`, `.ln_final`, `.text_projection`, `.logit_scale`, `.dtype`)."""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn


def _no_forward(self, *a, **k):
    raise RuntimeError("lecb200 parameter containers do not compute; use DenseCLIPB200 / the engine")


class _Holder(nn.Module):
    forward = _no_forward


def _bottleneck(inplanes, planes, stride):
    blk = _Holder()
    blk.conv1, blk.bn1 = nn.Conv2d(inplanes, planes, 1, bias=False), nn.BatchNorm2d(planes)
    blk.conv2, blk.bn2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False), nn.BatchNorm2d(planes)
    blk.conv3, blk.bn3 = nn.Conv2d(planes, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4)
    blk.stride = stride
    blk.downsample = None
    if stride > 1 or inplanes != planes * 4:
        blk.downsample = nn.Sequential(OrderedDict([
            ("-1", nn.AvgPool2d(stride)), ("0", nn.Conv2d(inplanes, planes * 4, 1, bias=False)),
            ("1", nn.BatchNorm2d(planes * 4))]))
    return blk


class ModifiedResNetParams(_Holder):
    """Weights of CLIP's ModifiedResNet (clip/model.py:130-190): 3-conv stem, 4 bottleneck stages, attnpool."""

    def __init__(self, layers, output_dim, heads, input_resolution=224, width=64):
        super().__init__()
        self.layers_cfg, self.output_dim, self.heads = tuple(layers), output_dim, heads
        self.input_resolution, self.width = input_resolution, width
        self.conv1, self.bn1 = nn.Conv2d(3, width // 2, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(width // 2)
        self.conv2, self.bn2 = nn.Conv2d(width // 2, width // 2, 3, padding=1, bias=False), nn.BatchNorm2d(width // 2)
        self.conv3, self.bn3 = nn.Conv2d(width // 2, width, 3, padding=1, bias=False), nn.BatchNorm2d(width)
        inplanes = width
        for i, (mult, blocks) in enumerate(zip((1, 2, 4, 8), layers), start=1):
            planes = width * mult
            stage = []
            for b in range(blocks):
                stage.append(_bottleneck(inplanes, planes, 2 if (i > 1 and b == 0) else 1))
                inplanes = planes * 4
            setattr(self, f"layer{i}", nn.Sequential(*stage))
        pool = _Holder()
        embed = width * 32
        pool.positional_embedding = nn.Parameter(torch.randn((input_resolution // 32) ** 2 + 1, embed) / embed ** 0.5)
        pool.k_proj, pool.q_proj, pool.v_proj = nn.Linear(embed, embed), nn.Linear(embed, embed), nn.Linear(embed, embed)
        pool.c_proj = nn.Linear(embed, output_dim)
        pool.num_heads = heads
        self.attnpool = pool


def _res_block(width, heads):
    blk = _Holder()
    blk.attn = nn.MultiheadAttention(width, heads)
    blk.ln_1 = nn.LayerNorm(width)
    blk.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(width, width * 4)), ("gelu", nn.Identity()),
                                         ("c_proj", nn.Linear(width * 4, width))]))
    blk.ln_2 = nn.LayerNorm(width)
    return blk


class TransformerParams(_Holder):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers, self.heads = width, layers, heads
        self.resblocks = nn.Sequential(*[_res_block(width, heads) for _ in range(layers)])


class VisionTransformerParams(_Holder):
    """Weights of CLIP's VisionTransformer (clip/model.py:240-276)."""

    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.input_resolution, self.output_dim, self.patch_size = input_resolution, output_dim, patch_size
        self.conv1 = nn.Conv2d(3, width, patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = TransformerParams(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))


class CLIPParams(_Holder):
    """Same constructor signature and state_dict keys as the reference `CLIP` (clip/model.py:279-333).
    ModifiedResNet towers are what the reference's dense path supports (T:365-373); VisionTransformer towers
    (BASELINE configs 3, 5) use this repo's dense definition (oracle/restatement.py `vit_dense`)."""

    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size,
                 context_length, vocab_size, transformer_width, transformer_heads, transformer_layers):
        super().__init__()
        self.context_length, self.vocab_size = context_length, vocab_size
        if isinstance(vision_layers, (tuple, list)):
            self.visual = ModifiedResNetParams(vision_layers, embed_dim, vision_width * 32 // 64, image_resolution, vision_width)
        else:
            self.visual = VisionTransformerParams(image_resolution, vision_patch_size, vision_width, vision_layers,
                                                  vision_width // 64, embed_dim)
        self.transformer = TransformerParams(transformer_width, transformer_layers, transformer_heads)
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width).normal_(std=0.01))
        self.ln_final = nn.LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim).normal_(std=transformer_width ** -0.5))
        self.logit_scale = nn.Parameter(torch.tensor(2.6592600))

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype


def describe(clip_model) -> dict:
    """Architecture facts the engine needs, read off any CLIP-shaped module (ours or the reference's)."""
    sd = clip_model.state_dict()
    tw = sd["ln_final.weight"].shape[0]
    if "visual.proj" in sd:                         # VisionTransformer tower
        width = sd["visual.conv1.weight"].shape[0]
        vit = dict(kind="vit", width=width, patch=sd["visual.conv1.weight"].shape[-1], vis_heads=width // 64,
                   layers=len({k.split(".")[3] for k in sd if k.startswith("visual.transformer.resblocks.")}),
                   tokens=sd["visual.positional_embedding"].shape[0])
    elif "visual.layer1.0.conv1.weight" in sd:
        width = sd["visual.layer1.0.conv1.weight"].shape[0]
        vit = dict(kind="rn", width=width, vis_heads=width * 32 // 64,
                   layers=tuple(len({k.split(".")[2] for k in sd if k.startswith(f"visual.layer{b}.")}) for b in (1, 2, 3, 4)))
    else:
        raise NotImplementedError("visual tower is neither a ModifiedResNet nor a VisionTransformer")
    return dict(**vit, embed_dim=sd["text_projection"].shape[1], text_width=tw, text_heads=tw // 64,
                text_layers=len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")}),
                context_length=sd["positional_embedding"].shape[0])
