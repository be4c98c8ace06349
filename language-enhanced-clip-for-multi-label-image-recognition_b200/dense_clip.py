"""Host-side mirror of the reference model API for the dual-prompt scoring path.

`DenseCLIPB200`, `PromptLearner`, `TextEncoder` keep the names, constructor arguments, forward
signatures, return arity/order, attribute names and state_dict keys of the reference classes in
project/my_code/trainers/Caption_distill_double.py (T) — `DenseCLIP` T:354-559, `PromptLearner`
T:104-308, `TextEncoder` T:72-101 — so the reference trainer (`build_model` T:755-760,
`model_inference` T:567-568, `forward_backward` T:804) can switch `cfg.TRAIN.MODEL` to this class and
existing prompt checkpoints (`ctx`, `ctx_double`, `ctx_evidence`, `temperature`, `spatial_T`,
`ranking_scale`, `token_prefix`, `token_suffix`, `token_suffix_nocls`) load unchanged.

All arithmetic runs in the sm_100a kernels of csrc/ through the C ABI; there is no PyTorch fallback.
Differences from the reference, all deliberate and listed in DESIGN.md:
  * the caption feature bank is an explicit constructor argument (the reference reads a module-level
    global loaded from a pickle at import, T:35-36); `None` skips retrieval (T:444-448);
  * the tokenizer is injectable (`tokenizer=` callable or pre-tokenised ids) because the reference's BPE
    vocabulary file is not part of this repo;
  * `reset_prompt_cache()` exposes the invalidation the reference lacks (T:421-439 never refreshes).
"""
from __future__ import annotations

import torch
from torch import nn

from . import checkpoint, ops
from .clip_model import describe
from .engine import TextTower, VisualRN, VisualViT


def _cfg_get(node, path, default=None):
    cur = node
    for part in path.split("."):
        try:
            cur = cur[part] if isinstance(cur, dict) else getattr(cur, part)
        except (KeyError, AttributeError):
            return default
    return cur


class TextEncoder(nn.Module):
    """T:72-101.  forward(prompts, tokenized_prompts, if_embedding=True, if_sequence=False)."""

    def __init__(self, clip_model):
        super().__init__()
        self.transformer = clip_model.transformer
        self.positional_embedding = clip_model.positional_embedding
        self.ln_final = clip_model.ln_final
        self.text_projection = clip_model.text_projection
        self.dtype = clip_model.dtype
        self.token_embedding = clip_model.token_embedding
        self._clip = [clip_model]           # not registered as a submodule twice
        self._tower = None

    def tower(self) -> TextTower:
        dev = self.positional_embedding.device
        if self._tower is None or self._tower.device != dev:
            if dev.type != "cuda":
                raise ops._lib.LecbError("lecb200 TextEncoder needs its weights on a CUDA device (no CPU path)")
            info = describe(self._clip[0])
            self._tower = TextTower(self._clip[0].state_dict(), info["text_width"], info["text_heads"],
                                    info["text_layers"], info["embed_dim"], dev)
        return self._tower

    @torch.no_grad()
    def forward(self, prompts, tokenized_prompts, if_embedding=True, if_sequence=False):
        tw = self.tower()
        if not if_embedding:
            tokenized_prompts = prompts
            prompts = tw.tok[prompts]                                   # token embedding gather (M:326)
        x = (prompts.float() + tw.pos).contiguous()
        if if_sequence:
            return tw.forward(x, sequence=True)
        eot = tokenized_prompts.argmax(dim=-1)
        # causal mask + EOT-row readout: positions after the last EOT cannot influence the result (exact trim)
        l_eff = int(eot.max()) + 1
        return tw.forward(x[:, :l_eff].contiguous(), eot_index=eot.to(x.device))


class PromptLearner(nn.Module):
    """T:104-308.  Parameters `ctx`, `ctx_double`, `ctx_evidence` ([n_ctx,W] or CSC [n_cls,n_ctx,W]),
    scalars `temperature`, `spatial_T`, `ranking_scale`; buffers `token_prefix`, `token_suffix`,
    `token_suffix_nocls`.  forward(neg_prompt_wcls=True) -> (prompts, prompts_neg, prompts_evidence,
    temperature, spatial_T, ranking_scale)."""

    def __init__(self, cfg, classnames, clip_model, nctx=None, tokenizer=None, tokenized_prompts=None,
                 tokenized_prompts_nocls=None):
        super().__init__()
        n_cls = len(classnames)
        n_ctx = _cfg_get(cfg, "TRAINER.Caption.N_CTX") if nctx is None else nctx
        ctx_init = _cfg_get(cfg, "TRAINER.Caption.CTX_INIT", "")
        csc = bool(_cfg_get(cfg, "TRAINER.Caption.CSC", False))
        dtype = clip_model.dtype
        ctx_dim = clip_model.ln_final.weight.shape[0]
        clip_imsize = clip_model.visual.input_resolution
        cfg_imsize = _cfg_get(cfg, "INPUT.SIZE")[0]
        assert cfg_imsize == clip_imsize, f"cfg_imsize ({cfg_imsize}) must equal to clip_imsize ({clip_imsize})"
        if tokenizer is None and tokenized_prompts is None:
            try:
                from clip import clip as _ref_clip          # running inside the reference tree
                tokenizer = lambda s: _ref_clip.tokenize(s, truncate=True)
            except Exception as e:  # pragma: no cover
                raise RuntimeError("PromptLearner needs `tokenizer=` (e.g. the reference clip.tokenize) or "
                                   "pre-tokenised `tokenized_prompts=`") from e
        if ctx_init:
            ctx_init = ctx_init.replace("_", " ")
            n_ctx = len(ctx_init.split(" "))
            with torch.no_grad():
                emb = clip_model.token_embedding(tokenizer(ctx_init).to(clip_model.token_embedding.weight.device)).type(dtype)
            vec = emb[0, 1:1 + n_ctx, :].detach().clone()
            ctx_vectors, ctx_vectors_double, ctx_vectors_evidence = vec, vec.clone(), vec.clone()
            prompt_prefix = ctx_init
        else:
            shape = (n_cls, n_ctx, ctx_dim) if csc else (n_ctx, ctx_dim)
            ctx_vectors = torch.empty(shape, dtype=dtype).normal_(std=0.02)
            ctx_vectors_double = torch.empty(shape, dtype=dtype).normal_(std=0.02)
            ctx_vectors_evidence = torch.empty((n_ctx, ctx_dim), dtype=dtype).normal_(std=0.02)   # never CSC (T:147)
            prompt_prefix = " ".join(["X"] * n_ctx)
        self.ctx = nn.Parameter(ctx_vectors)
        self.ctx_double = nn.Parameter(ctx_vectors_double)
        self.ctx_evidence = nn.Parameter(ctx_vectors_evidence)
        self.temperature = nn.Parameter(torch.tensor(3.0, dtype=dtype))
        self.spatial_T = nn.Parameter(torch.tensor(3.0, dtype=dtype))
        self.ranking_scale = nn.Parameter(torch.tensor(4.0, dtype=dtype))

        classnames = [name.replace("_", " ") for name in classnames]
        if tokenized_prompts is None:
            tokenized_prompts = torch.cat([tokenizer(prompt_prefix + " " + name + ".") for name in classnames])
        if tokenized_prompts_nocls is None:
            if tokenizer is not None:
                tokenized_prompts_nocls = torch.cat([tokenizer(prompt_prefix + ".") for _ in classnames])
            else:   # "X ... X ." == the class prompt with the class-name tokens removed: derive it
                tokenized_prompts_nocls = torch.zeros_like(tokenized_prompts)
                eot = tokenized_prompts.argmax(dim=-1)
                for i in range(n_cls):
                    head = tokenized_prompts[i, :1 + n_ctx]
                    tail = tokenized_prompts[i, eot[i] - 1:eot[i] + 1]          # ".", EOT
                    tokenized_prompts_nocls[i, :1 + n_ctx] = head
                    tokenized_prompts_nocls[i, 1 + n_ctx:3 + n_ctx] = tail
        assert tuple(tokenized_prompts.shape) == (n_cls, clip_model.positional_embedding.shape[0])
        dev = clip_model.token_embedding.weight.device
        with torch.no_grad():
            embedding = clip_model.token_embedding(tokenized_prompts.to(dev)).type(dtype)
            embedding_nocls = clip_model.token_embedding(tokenized_prompts_nocls.to(dev)).type(dtype)
        self.register_buffer("token_prefix", embedding[:, :1, :].clone())               # SOS
        self.register_buffer("token_suffix", embedding[:, 1 + n_ctx:, :].clone())       # CLS, EOS
        self.register_buffer("token_suffix_nocls", embedding_nocls[:, 1 + n_ctx:, :].clone())
        self.n_cls, self.n_ctx = n_cls, n_ctx
        self.tokenized_prompts = tokenized_prompts
        # T:173 `name_lens = [len(_tokenizer.encode(name)) ...]`: the tokens between the context and the final "."
        # (the BPE splits punctuation off first, so the class name tokenises the same inside the prompt)
        eot = tokenized_prompts.argmax(dim=-1)
        self.name_lens = [int(e) - 2 - n_ctx for e in eot]
        self.class_token_position = _cfg_get(cfg, "TRAINER.Caption.CLASS_TOKEN_POSITION", "end")

    def forward(self, neg_prompt_wcls=True):
        if self.class_token_position != "end":
            # the reference's "middle"/"front" branches never define prompts_neg and raise at T:308
            raise ValueError(f"CLASS_TOKEN_POSITION={self.class_token_position!r}: only 'end' is defined (T:217-308)")

        def expand(c):
            return c.unsqueeze(0).expand(self.n_cls, -1, -1) if c.dim() == 2 else c

        suffix_neg = self.token_suffix if neg_prompt_wcls else self.token_suffix_nocls
        prompts = torch.cat([self.token_prefix, expand(self.ctx), self.token_suffix], dim=1)
        prompts_neg = torch.cat([self.token_prefix, expand(self.ctx_double), suffix_neg], dim=1)
        prompts_evidence = torch.cat([self.token_prefix, expand(self.ctx_evidence), suffix_neg], dim=1)
        return prompts, prompts_neg, prompts_evidence, self.temperature, self.spatial_T, self.ranking_scale


class DenseCLIPB200(nn.Module):
    """Drop-in for `DenseCLIP` (T:354-559).

    forward(image=None, captions=None, if_test=False, model_name='ema'):
      test  : image [B,3,H,W] float -> (logits_ [B,K], logits_local [B,K], logits_neg [P,B,K],
              feats·T_posᵀ [P,B,K], topk_scores [B,10] (zeros without a caption bank))        (T:472)
              `image` may also be a uint8 NHWC batch [B,H,W,3] of raw pixels (normalised inside the stem kernel)
      train : captions [B,77] int64 -> (logits_, logits_local, image_features [L,B,D],
              text_features [K,D], logits_m_ | None, logits_local_m | None)                  (T:545)
    """

    def __init__(self, cfg, classnames, clip_model, return_interm_layers=False, nctx=None, caption_bank=None,
                 tokenizer=None, tokenized_prompts=None):
        super().__init__()
        kw = dict(tokenizer=tokenizer, tokenized_prompts=tokenized_prompts)
        self.prompt_learner = PromptLearner(cfg, classnames, clip_model, nctx, **kw)
        # checkpoint.load_model() / load_pretrained_weights() receive the prompt learner alone (the trainer registers it,
        # T:774) and must drop this module's cached text features: a weak back-reference, invisible to nn.Module
        checkpoint.register_owner(self.prompt_learner, self)
        # Opt-in (not in the reference): True or a process group splits the 2-3 x K prompt sequences of
        # the prompt-tuning step over the ranks instead of replicating them (train_path.py, dist.py); gradients after
        # dist.allreduce_mean_grads are those of the replicated branch
        self.shard_prompt_branch = False
        self.prompt_learner_m = PromptLearner(cfg, classnames, clip_model, nctx, **kw)
        self.tokenized_prompts = self.prompt_learner.tokenized_prompts
        self.text_encoder = TextEncoder(clip_model)
        self.model = clip_model
        self.return_interm_layers = return_interm_layers
        ap = getattr(clip_model.visual, "attnpool", None)
        if ap is not None:                     # ModifiedResNet tower (the reference's only dense path)
            # T:364-368: `visual_encoder = IntermediateLayerGetter(self.model.visual, return_layers)` re-registers the trunk's
            # children (up to the last returned layer) under a second name, and a slice of the attnpool positional embedding is
            # kept as a plain attribute: no compute here, but the same state_dict keys / attribute surface as the reference
            last = "layer4"
            alias = nn.ModuleDict()
            for name, child in clip_model.visual.named_children():
                alias[name] = child
                if name == last:
                    break
            self.visual_encoder = alias
            self.positional_embedding = ap.positional_embedding[1::]
            self.v_linear_weight, self.v_linear_bias = ap.v_proj.weight, ap.v_proj.bias      # aliases, T:370-373
            self.c_linear_weight, self.c_linear_bias = ap.c_proj.weight, ap.c_proj.bias
        self.logit_scale = clip_model.logit_scale
        self.dtype = clip_model.dtype
        self.cfg = cfg
        self.caption_bank = caption_bank          # [N,D] fp16/fp32 tensor or None (replaces the T:35-36 global)
        self.prompt_text_features = None
        self.model_pairs = [[self.prompt_learner, self.prompt_learner_m]]
        self.copy_params()
        self._visual = None
        self._packed_text = None
        self._info = describe(clip_model)

    # ---------------------------------------------------------------- engines
    def visual_engine(self):
        dev = self.model.visual.conv1.weight.device
        if self._visual is None or self._visual.device != dev:
            if dev.type != "cuda":
                raise ops._lib.LecbError("lecb200 DenseCLIPB200 needs its weights on a CUDA device (no CPU path)")
            i = self._info
            if i["kind"] == "vit":
                self._visual = VisualViT(self.model.state_dict(), i["patch"], i["width"], i["layers"], i["vis_heads"],
                                         i["embed_dim"], dev)
            else:
                self._visual = VisualRN(self.model.state_dict(), i["layers"], i["width"], i["vis_heads"], i["embed_dim"], dev)
        return self._visual

    def reset_prompt_cache(self):
        """Drop the cached prompt features so the next test forward re-encodes the prompts."""
        self.prompt_text_features = None

    def encode_image(self, x):
        """T:385-399.  Returns NCHW fp32 like the reference (a view of the engine's NHWC bf16 output)."""
        if self._info["kind"] == "vit":
            raise NotImplementedError("encode_image returns the ModifiedResNet layer4 map (T:385-399); ViT towers have none")
        return self.visual_engine().trunk(x.float()).permute(0, 3, 1, 2).float()

    # ---------------------------------------------------------------- forward
    def forward(self, image=None, captions=None, if_test=False, model_name='ema'):
        if if_test:
            return self._forward_test(image)
        from .train_path import forward_train
        return forward_train(self, captions)

    def _scales(self, temperature, spatial_T, which):
        learn = bool(_cfg_get(self.cfg, "TRAIN.IF_LEARN_SCALE", False))
        learn_sp = bool(_cfg_get(self.cfg, "TRAIN.IF_LEARN_spatial_SCALE", False))
        logit_scale = float(temperature.exp()) if learn else 4.0
        spatial = float(spatial_T.exp()) if learn_sp else float(_cfg_get(self.cfg, f"TRAIN.spatial_SCALE_{which}"))
        return logit_scale, spatial

    @torch.no_grad()
    def _prompt_features(self, use_evidence):
        """T:421-439: encode the K prompts once, L2-normalise, cache."""
        if self.prompt_text_features is None:
            prompts, prompts_double, prompts_evidence, _, _, _ = self.prompt_learner()
            tok = self.tokenized_prompts
            feats = {"text_features": ops.l2norm_rows(self.text_encoder(prompts, tok)),
                     "text_features_neg": ops.l2norm_rows(self.text_encoder(prompts_double, tok))}
            if use_evidence:
                feats["text_features_evidence"] = ops.l2norm_rows(self.text_encoder(prompts_evidence, tok))
            self.prompt_text_features = feats
        return self.prompt_text_features

    @torch.no_grad()
    def _forward_test(self, image):
        use_evidence = bool(_cfg_get(self.cfg, "TRAINER.Caption.use_evidence", False))
        eng = self.visual_engine()
        row_mask = None
        if self._info["kind"] == "vit":
            # token rows [B, T]: row 0 of every image is the class token = the global feature; it is masked out of
            # the spatial aggregation exactly like a padded caption token (lecb_head_aggregate row_mask)
            local, ssq, p = eng.tokens(image.float())
            b = image.shape[0]
            g = local.view(b, p, -1)[:, 0].float().contiguous()
            if getattr(self, "_cls_mask", None) is None or self._cls_mask.shape[0] != b * p or self._cls_mask.device != local.device:
                m = torch.zeros((b, p), device=local.device, dtype=torch.uint8)
                m[:, 0] = 1
                self._cls_mask = m.view(-1)
            row_mask = self._cls_mask
        else:
            # additive: a uint8 NHWC batch [B,H,W,3] of raw pixels (what lecb_crop_resize_u8 produces) is normalised inside
            # the stem kernel with cfg.INPUT.PIXEL_MEAN / PIXEL_STD; a float batch is the reference's normalised NCHW tensor
            if image.dtype == torch.uint8:
                feat = eng.trunk(image, mean=_cfg_get(self.cfg, "INPUT.PIXEL_MEAN", ops.CLIP_PIXEL_MEAN),
                                 std=_cfg_get(self.cfg, "INPUT.PIXEL_STD", ops.CLIP_PIXEL_STD))
            else:
                feat = eng.trunk(image.float())
            b, h, w, _ = feat.shape
            p = h * w
            local, ssq, g = eng.pooled(feat)
        _, _, _, temperature, spatial_T, _ = self.prompt_learner()
        tf = self._prompt_features(use_evidence)
        t_pos, t_neg = tf["text_features"], tf["text_features_neg"]
        k = t_pos.shape[0]
        names = ["text_features", "text_features_neg"] + (["text_features_evidence"] if use_evidence else [])
        if self._packed_text is None or self._packed_text[0] is not tf or self._packed_text[1] != tuple(names):
            cat = torch.cat([tf[n] for n in names], 0)
            pad = (-cat.shape[0]) % 8
            if pad:
                cat = torch.cat([cat, cat.new_zeros((pad, cat.shape[1]))], 0)
            self._packed_text = (tf, tuple(names), cat.to(torch.bfloat16).contiguous())
        logit_scale, spatial = self._scales(temperature, spatial_T, "image")
        dots = ops.gemm(local, self._packed_text[2], out_f32=True)                      # [B*P, n_txt*K] raw dot products
        logits_local, neg_map, pos_map = ops.head_aggregate(dots, b, p, k, len(names), row_sumsq=ssq, row_mask=row_mask,
                                                            logit_scale=logit_scale, spatial_scale=spatial)
        if row_mask is not None:
            neg_map, pos_map = neg_map[1:], pos_map[1:]          # drop the class-token row: [P,B,K] patch maps
        g_unit = ops.l2norm_rows(g)
        g_add = None
        if self.caption_bank is not None:
            from .retrieval import retrieve_mean
            g_add, topk_scores = retrieve_mean(g_unit, self.caption_bank)
        else:
            # no bank: the global feature is used as is; the 5th return value keeps the reference's shape so that the
            # trainer's `sim.reshape(...)` / `.cpu()` (T:645, T:701) works unchanged
            topk_scores = torch.zeros((b, 10), device=g_unit.device, dtype=torch.float32)
        logits_ = ops.global_logits(g_unit, t_pos, g_add, logit_scale)
        return logits_, logits_local, neg_map, pos_map, topk_scores

    # ---------------------------------------------------------------- EMA twin (T:547-559)
    @torch.no_grad()
    def copy_params(self):
        for live, twin in self.model_pairs:
            for p, pm in zip(live.parameters(), twin.parameters()):
                pm.data.copy_(p.data)
                pm.requires_grad = False

    @torch.no_grad()
    def _momentum_update(self):
        """T:554-559, one multi-tensor launch (lecb_ema_update) instead of three ATen kernels per parameter."""
        m = float(_cfg_get(self.cfg, "TRAIN.momentum", 0.995))
        for live, twin in self.model_pairs:
            lp, tp = [p.data for p in live.parameters()], [p.data for p in twin.parameters()]
            if all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in lp + tp):
                ops.ema_update(lp, tp, m)
            else:
                raise ops._lib.LecbError("lecb200: the EMA twin update needs fp32 CUDA prompt parameters (no CPU path)")


DenseCLIP = DenseCLIPB200   # the reference's class name, for `cfg.TRAIN.MODEL == "DenseCLIP"` call sites
