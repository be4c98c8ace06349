"""Host mirror of the adapter variant of the model (reference: trainers/Caption_distill_double_adapter.py, TA):
`Adapter` TA:304-317, `AdapterTextEncoder` TA:86-125, `PromptLearner` TA:127-318 (two contexts, no evidence prompt, 5-tuple),
`AdapterDenseCLIP` TA:320-457 (4-tuple returns, no EMA twin, no caption retrieval).

What differs from `DenseCLIPB200`: the PROMPTS go through `AdapterTextEncoder` — the CLIP text transformer, then
`x + Adapter(x)` (512 -> 128 -> 512, two bias-free linears, ReLU after each) before `ln_final` — while the captions keep the
plain `TextEncoder`.  The adapter's weights are part of the state_dict but FROZEN in the reference: its `build_model` switches
off every parameter whose name lacks "prompt_learner" (TA:534-536) and hands only the prompt learner to the optimiser (TA:544),
so the backward needs the adapter's data gradient only.  The reference re-encodes the prompts on every test forward; here they
are cached (`reset_prompt_cache()` invalidates), like `DenseCLIPB200`.
All arithmetic runs in the sm_100a kernels of csrc/ (engine.TextTower with `set_adapter`); there is no PyTorch fallback."""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .clip_model import describe
from .dense_clip import PromptLearner as _FullPromptLearner
from .dense_clip import TextEncoder, _cfg_get
from .engine import TextTower, VisualRN


class Adapter(nn.Module):
    """TA:304-317 (parameter container: `fc.0.weight` [c_in/r, c_in], `fc.2.weight` [c_in, c_in/r])."""

    def __init__(self, c_in, reduction=4):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c_in, c_in // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(c_in // reduction, c_in, bias=False), nn.ReLU(inplace=True))

    def forward(self, x):
        raise ops._lib.LecbError("lecb200 Adapter is a parameter container: it runs inside AdapterTextEncoder (no PyTorch path)")


class AdapterTextEncoder(TextEncoder):
    """TA:86-125.  Same signature as `TextEncoder.forward`; prompts only (the reference calls it with embeddings + EOT readout)."""

    def __init__(self, clip_model):
        super().__init__(clip_model)
        self.text_adapter = Adapter(512, 4).to(clip_model.dtype)           # TA:97 (the literal 512 is the reference's)

    def tower(self) -> TextTower:
        fresh = self._tower is None or self._tower.device != self.positional_embedding.device
        tw = super().tower()
        if fresh or tw.adapter is None:
            tw.set_adapter(self.text_adapter.fc[0].weight, self.text_adapter.fc[2].weight)
        return tw

    def refresh_adapter(self):
        """Re-read the adapter weights (after load_state_dict)."""
        if self._tower is not None:
            self._tower.set_adapter(self.text_adapter.fc[0].weight, self.text_adapter.fc[2].weight)

    @torch.no_grad()
    def forward(self, prompts, tokenized_prompts, if_embedding=True, if_sequence=False):
        if if_sequence:
            raise ops._lib.LecbError("AdapterTextEncoder is used for prompts only (TA:375-376, 418-419)")
        return super().forward(prompts, tokenized_prompts, if_embedding=if_embedding, if_sequence=False)


class PromptLearner(_FullPromptLearner):
    """TA:127-318: `ctx`, `ctx_double`, the three scalars and the three token buffers; forward -> 5-tuple."""

    def __init__(self, cfg, classnames, clip_model, tokenizer=None, tokenized_prompts=None, tokenized_prompts_nocls=None):
        super().__init__(cfg, classnames, clip_model, None, tokenizer=tokenizer, tokenized_prompts=tokenized_prompts,
                         tokenized_prompts_nocls=tokenized_prompts_nocls)
        del self.ctx_evidence                                               # the adapter trainer's learner has no evidence context

    def forward(self, neg_prompt_wcls=True):
        if self.class_token_position != "end":
            raise ValueError(f"CLASS_TOKEN_POSITION={self.class_token_position!r}: only 'end' defines prompts_neg (TA:224-300)")

        def expand(c):
            return c.unsqueeze(0).expand(self.n_cls, -1, -1) if c.dim() == 2 else c

        suffix_neg = self.token_suffix if neg_prompt_wcls else self.token_suffix_nocls
        prompts = torch.cat([self.token_prefix, expand(self.ctx), self.token_suffix], dim=1)
        prompts_neg = torch.cat([self.token_prefix, expand(self.ctx_double), suffix_neg], dim=1)
        return prompts, prompts_neg, self.temperature, self.spatial_T, self.ranking_scale


class AdapterDenseCLIPB200(nn.Module):
    """Drop-in for `AdapterDenseCLIP` (TA:320-457).

    forward(image=None, captions=None, if_test=False):
      test  : image [B,3,H,W] float (or uint8 NHWC) -> (logits_ [B,K], logits_local [B,K], logits_neg [P,B,K], feats·T_posᵀ [P,B,K])
      train : captions [B,77] int64 -> (logits_, logits_local, image_features [L,B,D], text_features [K,D])"""

    def __init__(self, cfg, classnames, clip_model, return_interm_layers=False, tokenizer=None, tokenized_prompts=None):
        super().__init__()
        self.prompt_learner = PromptLearner(cfg, classnames, clip_model, tokenizer=tokenizer, tokenized_prompts=tokenized_prompts)
        self.tokenized_prompts = self.prompt_learner.tokenized_prompts
        self.text_encoder = TextEncoder(clip_model)
        self.adapter_text_encoder = AdapterTextEncoder(clip_model)
        self.model = clip_model
        self.return_interm_layers = return_interm_layers
        ap = clip_model.visual.attnpool
        alias = nn.ModuleDict()                                            # IntermediateLayerGetter aliases (TA:339)
        for name, child in clip_model.visual.named_children():
            alias[name] = child
            if name == "layer4":
                break
        self.visual_encoder = alias
        self.positional_embedding = ap.positional_embedding[1::]
        self.v_linear_weight, self.v_linear_bias = ap.v_proj.weight, ap.v_proj.bias
        self.c_linear_weight, self.c_linear_bias = ap.c_proj.weight, ap.c_proj.bias
        self.logit_scale = clip_model.logit_scale
        self.dtype = clip_model.dtype
        self.cfg = cfg
        self.prompt_text_features = None
        self._visual = None
        self._packed_text = None
        self._info = describe(clip_model)

    def visual_engine(self):
        dev = self.model.visual.conv1.weight.device
        if self._visual is None or self._visual.device != dev:
            if dev.type != "cuda":
                raise ops._lib.LecbError("lecb200 AdapterDenseCLIPB200 needs its weights on a CUDA device (no CPU path)")
            i = self._info
            self._visual = VisualRN(self.model.state_dict(), i["layers"], i["width"], i["vis_heads"], i["embed_dim"], dev)
        return self._visual

    def reset_prompt_cache(self):
        self.prompt_text_features = None

    def forward(self, image=None, captions=None, if_test=False):
        if if_test:
            return self._forward_test(image)
        return self._forward_train(captions)

    @torch.no_grad()
    def _forward_test(self, image):
        eng = self.visual_engine()
        if image.dtype == torch.uint8:
            feat = eng.trunk(image, mean=_cfg_get(self.cfg, "INPUT.PIXEL_MEAN", ops.CLIP_PIXEL_MEAN),
                             std=_cfg_get(self.cfg, "INPUT.PIXEL_STD", ops.CLIP_PIXEL_STD))
        else:
            feat = eng.trunk(image.float())
        b, h, w, _ = feat.shape
        p = h * w
        local, ssq, g = eng.pooled(feat)
        prompts, prompts_double, temperature, spatial_T, _ = self.prompt_learner()
        if self.prompt_text_features is None:
            tok = self.tokenized_prompts
            self.prompt_text_features = {"text_features": ops.l2norm_rows(self.adapter_text_encoder(prompts, tok)),
                                         "text_features_neg": ops.l2norm_rows(self.adapter_text_encoder(prompts_double, tok))}
            self._packed_text = None
        tf = self.prompt_text_features
        k = tf["text_features"].shape[0]
        if self._packed_text is None:
            cat = torch.cat([tf["text_features"], tf["text_features_neg"]], 0)
            pad = (-cat.shape[0]) % 8
            if pad:
                cat = torch.cat([cat, cat.new_zeros((pad, cat.shape[1]))], 0)
            self._packed_text = cat.to(torch.bfloat16).contiguous()
        learn = bool(_cfg_get(self.cfg, "TRAIN.IF_LEARN_SCALE", False))
        learn_sp = bool(_cfg_get(self.cfg, "TRAIN.IF_LEARN_spatial_SCALE", False))
        logit_scale = float(temperature.exp()) if learn else 4.0
        spatial = float(spatial_T.exp()) if learn_sp else float(_cfg_get(self.cfg, "TRAIN.spatial_SCALE_image"))
        dots = ops.gemm(local, self._packed_text, out_f32=True)
        logits_local, neg_map, pos_map = ops.head_aggregate(dots, b, p, k, 2, row_sumsq=ssq, logit_scale=logit_scale,
                                                            spatial_scale=spatial)
        logits_ = ops.global_logits(ops.l2norm_rows(g), tf["text_features"], None, logit_scale)
        return logits_, logits_local, neg_map, pos_map

    def _forward_train(self, captions):
        from . import train_path as TP
        if bool(_cfg_get(self.cfg, "TRAIN.IF_LEARN_spatial_SCALE", False)):
            raise NotImplementedError("lecb200: a learnable spatial scale is not supported on the prompt-tuning path")
        l_full = captions.shape[1]
        l_run = TP.caption_run_length(self, captions)
        captions = captions.to(self.text_encoder.positional_embedding.device)
        b = captions.shape[0]
        local, ssq, mask, g_unit = TP._caption_branch(self, captions, l_run)
        prompts, prompts_double, temperature, spatial_T, _ = self.prompt_learner()
        learn = bool(_cfg_get(self.cfg, "TRAIN.IF_LEARN_SCALE", False))
        logit_scale = float(temperature.exp()) if learn else 4.0
        spatial = float(_cfg_get(self.cfg, "TRAIN.spatial_SCALE_text"))
        if getattr(self, "_eot_dev", None) is None or self._eot_dev[0].device != local.device:
            eot = self.tokenized_prompts.argmax(dim=-1)
            self._eot_dev = (eot.to(local.device), int(eot.max()) + 1)
        pack = (self.adapter_text_encoder.tower(), self._eot_dev, local, ssq, mask, g_unit, b, l_run, logit_scale, spatial, None)
        logits, logits_local, text_features = TP._DualPromptHead.apply(pack, temperature if learn else None, prompts, prompts_double)
        with torch.no_grad():
            feats = ops.l2norm_rows(local, out_dtype=torch.float32).view(b, l_run, -1)
            if l_run < l_full:
                feats = torch.nn.functional.pad(feats, (0, 0, 0, l_full - l_run))
        return logits, logits_local, feats.permute(1, 0, 2), text_features


AdapterDenseCLIP = AdapterDenseCLIPB200   # the reference's class name
