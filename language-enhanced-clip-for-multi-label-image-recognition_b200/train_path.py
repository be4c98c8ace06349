"""Text-only prompt-tuning forward/backward (reference: DenseCLIP.forward(None, captions), T:473-545).

The caption branch (text-as-image features) is frozen and runs without saving anything.  The prompt
branch — 2 or 3 x K prompt sequences through the text tower, L2 normalisation, the caption-by-class
contraction, winner-take-all + token-softmax aggregation and the global logits — is one
torch.autograd.Function whose backward is hand-written kernels end to end; torch autograd only carries the
gradient from the [K,77,W] prompt embeddings back into `ctx` / `ctx_double` / `ctx_evidence` through the
reference's own `torch.cat` / `expand` in PromptLearner.forward (T:199-242)."""
from __future__ import annotations

import torch

from . import dist, ops


def _cfg(model, path, default=None):
    from .dense_clip import _cfg_get
    return _cfg_get(model.cfg, path, default)


class _DualPromptHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pack, logit_scale_t, *prompts):
        tower, (tok_prompts, l_eff), local, ssq, mask, g_unit, b, l, logit_scale, spatial, shard_group = pack
        rider = None
        if isinstance(local, dict):              # the caption branch has not run yet: it rides along with the prompt rows
            rider = local
        n_txt = len(prompts)
        k = prompts[0].shape[0]
        # The mask is causal (M:364-370) and only the EOT row of a prompt is used (T:100), so positions after the
        # last EOT can never reach the output or receive gradient: run the tower on the first l_eff positions only
        # (exact, ~3x less work for "X*16 <class>." prompts whose EOT sits near position 20 of 77).
        ctx.full_len = prompts[0].shape[1]
        x = (torch.cat([p.detach().float() for p in prompts], 0) + tower.pos)[:, :l_eff].contiguous()   # [n*K, l_eff, W]
        eot = tok_prompts.repeat(n_txt)          # EOT index per prompt sequence, already on the device
        ctx.rows = None
        xr, er, group = x, eot, None
        if shard_group is not None:
            # class-sharded prompt branch (dist.py): the tower runs on this rank's rows, the features are all-gathered
            group = None if shard_group is True else shard_group
            lo, hi, _ = dist.my_chunk(x.shape[0], group)
            if hi <= lo:
                raise RuntimeError(f"shard_prompt_branch: {x.shape[0]} prompt sequences cannot feed every rank")
            xr, er = x[lo:hi].contiguous(), eot[lo:hi].contiguous()
            ctx.rows = (lo, hi, group, x.shape[0], x.shape[1], x.shape[2])
        if rider is not None:
            t_raw, saved, tail = tower.forward_train(xr, er, rider=rider["x"])
            local, ssq, mask, g_unit = _caption_tail(tower, tail, rider["captions"])
            rider["out"] = (local, ssq, mask, g_unit)
        else:
            t_raw, saved = tower.forward_train(xr, er)                                            # [n*K, D] fp32
        if shard_group is not None:
            t_raw = dist.gather_rows(t_raw, x.shape[0], group)
        t_hat = ops.l2norm_rows(t_raw)
        pad = (-t_hat.shape[0]) % 8
        t_cat = t_hat if not pad else torch.cat([t_hat, t_hat.new_zeros((pad, t_hat.shape[1]))], 0)
        dots = ops.gemm(local, t_cat.to(torch.bfloat16).contiguous(), out_f32=True)               # [B*L, n*K(+pad)]
        logits_local, _, _ = ops.head_aggregate(dots, b, l, k, n_txt, row_sumsq=ssq, row_mask=mask,
                                                logit_scale=logit_scale, spatial_scale=spatial, want_maps=False)
        t_pos = t_hat[:k].contiguous()
        logits = ops.global_logits(g_unit, t_pos, None, logit_scale)
        ctx.pack = (tower, saved, t_raw, dots, local, ssq, mask, g_unit, b, l, k, n_txt, logit_scale, spatial)
        ctx.save_for_backward(logits, logits_local)
        ctx.learn_scale = logit_scale_t is not None and logit_scale_t.requires_grad
        return logits, logits_local, t_pos

    @staticmethod
    def backward(ctx, d_logits, d_local, d_tpos):
        tower, saved, t_raw, dots, local, ssq, mask, g_unit, b, l, k, n_txt, logit_scale, spatial = ctx.pack
        logits, logits_local = ctx.saved_tensors
        dev = dots.device
        d_logits = torch.zeros((b, k), device=dev) if d_logits is None else d_logits.float().contiguous()
        d_local = torch.zeros((b, k), device=dev) if d_local is None else d_local.float().contiguous()
        d_dots = ops.head_aggregate_bwd(dots, d_local, b, l, k, n_txt, row_sumsq=ssq, row_mask=mask,
                                        logit_scale=logit_scale, spatial_scale=spatial)
        d_that = ops.tn_gemm_small(d_dots, local, n_txt * k)                                       # [n*K, D]
        ops.tn_gemm_small(d_logits, g_unit, k, out=d_that, alpha=logit_scale, accumulate=True)     # rows [0,K) = positive prompts
        if d_tpos is not None:
            d_that[:k] += d_tpos.float()
        d_traw = ops.l2norm_bwd(t_raw, d_that)
        if ctx.rows is None:
            dx = tower.backward(saved, d_traw)                                                      # [n*K, l_eff, W]
        else:
            # every rank's loss depends on every text feature: sum the feature gradients over ranks, then back-propagate
            # this rank's rows only (the other rows of dx stay zero; the flat gradient average completes the sum)
            lo, hi, group, n_all, l_run, w_run = ctx.rows
            d_traw = dist.sum_over_ranks(d_traw.contiguous(), group)
            dx_own = tower.backward(saved, d_traw[lo:hi].contiguous())
            dx = dx_own.new_zeros((n_all, l_run, w_run))
            dx[lo:hi] = dx_own
        if dx.shape[1] < ctx.full_len:           # positions past the last EOT carry exactly zero gradient
            dx = torch.nn.functional.pad(dx, (0, 0, 0, ctx.full_len - dx.shape[1]))
        grads = tuple(dx[i * k:(i + 1) * k] for i in range(n_txt))
        d_scale = None
        if ctx.learn_scale:          # logits = exp(temperature) * (...)  =>  dL/dtemperature = sum(dlogits * logits)
            d_scale = (d_logits * logits).sum() + (d_local * logits_local).sum()
        ctx.pack = None
        return (None, d_scale) + grads


def caption_run_length(model, captions):
    """Positions of the caption tower that can matter: everything up to the last EOT of the batch.  The mask is causal
    (M:364-370), so later positions cannot influence earlier ones, and they are padding (token id 0) whose logits get -10000
    before the token softmax (T:491-498): their weight is exactly zero in fp32.  Real captions end near position 20 of 77
    (74 % of the caption FLOPs are padding), so the tower runs on [B, l_run] instead of [B, 77].
    Needs one host read of the batch's longest caption; under CUDA-graph capture (no host reads) the caller's promise
    `model.caption_len_hint` is used, else the full length."""
    l = captions.shape[1]
    if not bool(getattr(model, "trim_caption_padding", True)):
        return l
    if captions.is_cuda and torch.cuda.is_current_stream_capturing():
        hint = getattr(model, "caption_len_hint", None)
        return l if hint is None else max(1, min(l, int(hint)))
    return max(1, min(l, int(captions.argmax(dim=-1).max()) + 1))


def _caption_inputs(model, captions, l_run=None):
    """T:474-475: token embedding + positional embedding of the first l_run caption positions -> (fp32 [B,l,W], captions[:, :l])."""
    tw = model.text_encoder.tower()
    if l_run is not None and l_run < captions.shape[1]:
        captions = captions[:, :l_run].contiguous()
    l = captions.shape[1]
    return (tw.tok[captions] + tw.pos[:l]).contiguous(), captions


def _caption_tail(tw, xs, captions):
    """T:476-477,485-486,491 after the transformer: ln_final, projection (+ row sum of squares), the EOT row as the global feature,
    the padding mask.  xs fp32 [B*l, W]."""
    b, l = captions.shape
    h, _, _, _ = ops.layernorm(xs, *tw.ln_final)
    ssq = torch.zeros((b * l,), device=xs.device, dtype=torch.float32)
    local = ops.gemm(h, tw.text_proj_t, row_sumsq=ssq)                       # bf16 [B*L, D] + row sum of squares
    eot = captions.argmax(dim=-1)
    g = local.view(b, l, -1)[torch.arange(b, device=xs.device), eot].float().contiguous()
    g_unit = ops.l2norm_rows(g)
    mask = (captions == 0).to(torch.uint8).reshape(-1).contiguous()          # token id 0 = padding (T:491)
    return local, ssq, mask, g_unit


@torch.no_grad()
def _caption_branch(model, captions, l_run=None):
    """T:474-477,485-486,491: per-token caption features (frozen path) for the first l_run positions."""
    tw = model.text_encoder.tower()
    x, captions = _caption_inputs(model, captions, l_run)
    n, l, w = x.shape
    xs = x.reshape(n * l, w)
    for blk in tw.blocks:
        h, _, _, _ = ops.layernorm(xs, *blk["ln1"])
        qkv = ops.gemm(h, *blk["qkv"])
        a = ops.causal_attn(qkv, n, l, w, tw.heads)
        xs = ops.gemm_f32res(a, *blk["out"], xs)
        h, _, _, _ = ops.layernorm(xs, *blk["ln2"])
        u = ops.gemm(h, *blk["fc"], quick_gelu=True)
        xs = ops.gemm_f32res(u, *blk["proj"], xs)
    return _caption_tail(tw, xs, captions)


def _prompt_logits_nograd(model, learner, local, ssq, mask, g_unit, b, l, logit_scale, spatial, use_evidence):
    """EMA twin branch (T:516-541): same math, no saves."""
    prompts, prompts_double, prompts_evidence, _, _, _ = learner()
    tok = model.tokenized_prompts
    feats = [ops.l2norm_rows(model.text_encoder(p, tok)) for p in
             ((prompts, prompts_double, prompts_evidence) if use_evidence else (prompts, prompts_double))]
    k = feats[0].shape[0]
    cat = torch.cat(feats, 0)
    pad = (-cat.shape[0]) % 8
    if pad:
        cat = torch.cat([cat, cat.new_zeros((pad, cat.shape[1]))], 0)
    dots = ops.gemm(local, cat.to(torch.bfloat16).contiguous(), out_f32=True)
    logits_local, _, _ = ops.head_aggregate(dots, b, l, k, len(feats), row_sumsq=ssq, row_mask=mask,
                                            logit_scale=logit_scale, spatial_scale=spatial, want_maps=False)
    return ops.global_logits(g_unit, feats[0], None, logit_scale), logits_local


def forward_train(model, captions):
    """-> (logits_, logits_local, image_features [L,B,D], text_features [K,D], logits_m_|None, logits_local_m|None)."""
    use_evidence = bool(_cfg(model, "TRAINER.Caption.use_evidence", False))
    if bool(_cfg(model, "TRAIN.IF_LEARN_spatial_SCALE", False)):
        raise NotImplementedError("lecb200: a learnable spatial scale is not supported on the prompt-tuning path "
                                  "(every shipped config fixes TRAIN.spatial_SCALE_text)")
    l_full = captions.shape[1]
    l_run = caption_run_length(model, captions)           # before the upload when the batch still lives on the host
    captions = captions.to(model.text_encoder.positional_embedding.device)
    b = captions.shape[0]
    l = l_run
    prompts, prompts_double, prompts_evidence, temperature, spatial_T, _ = model.prompt_learner()
    learn = bool(_cfg(model, "TRAIN.IF_LEARN_SCALE", False))
    logit_scale = float(temperature.detach().exp()) if learn else 4.0
    spatial = float(_cfg(model, "TRAIN.spatial_SCALE_text"))
    if getattr(model, "_eot_dev", None) is None or model._eot_dev[0].device != captions.device:
        eot = model.tokenized_prompts.argmax(dim=-1)
        # cached (EOT indices on the device, live prompt length): keeps the step free of host syncs / graph-capturable
        model._eot_dev = (eot.to(captions.device), int(eot.max()) + 1)
    shard = getattr(model, "shard_prompt_branch", False)
    if shard and not dist.multi_rank(None if shard is True else shard):
        shard = False                    # single process: nothing to shard over
    if shard:
        # feasibility is decided from values identical on every rank BEFORE any collective: with ceil(n / world)-row chunks
        # the last ranks can end up with no rows (40 sequences on 16 ranks), and a rank that raised on its own would
        # leave the others hanging in the all-gather; such a layout simply keeps the replicated branch everywhere
        import torch.distributed as _d
        world = _d.get_world_size(None if shard is True else shard)
        n_rows = (3 if use_evidence else 2) * prompts.shape[0]
        if -(-n_rows // world) * (world - 1) >= n_rows:
            shard = False
    plist = (prompts, prompts_double, prompts_evidence) if use_evidence else (prompts, prompts_double)
    # One tower pass for both branches (the frozen caption rows ride along with the prompt rows — this rank's, when the prompt
    # branch is sharded — through every row-wise launch: TextTower.forward_train) unless the model asks for two passes.
    joint = bool(getattr(model, "joint_text_pass", True)) and torch.is_grad_enabled() and any(p.requires_grad for p in plist)
    if joint:
        with torch.no_grad():
            cx, ccaps = _caption_inputs(model, captions, l_run)
        rider = {"x": cx, "captions": ccaps}
        pack = (model.text_encoder.tower(), model._eot_dev, rider, None, None, None, b, l, logit_scale, spatial,
                (True if shard is True else shard) if shard else None)
    else:
        local, ssq, mask, g_unit = _caption_branch(model, captions, l_run)
        pack = (model.text_encoder.tower(), model._eot_dev, local, ssq, mask, g_unit, b, l, logit_scale, spatial,
                (True if shard is True else shard) if shard else None)
    logits, logits_local, text_features = _DualPromptHead.apply(pack, temperature if learn else None, *plist)
    if joint:
        local, ssq, mask, g_unit = rider.pop("out")
    with torch.no_grad():
        feats = ops.l2norm_rows(local, out_dtype=torch.float32).view(b, l, -1)
        if l < l_full:
            # positions past the batch's last EOT were not run: zeros there (the reference holds the normalised features of
            # pad tokens, which its own -10000 mask keeps from reaching any logit)
            feats = torch.nn.functional.pad(feats, (0, 0, 0, l_full - l))
        image_features = feats.permute(1, 0, 2)
    logits_m, logits_local_m = None, None
    if bool(_cfg(model, "TRAIN.ema", False)):
        with torch.no_grad():
            model._momentum_update()
            logits_m, logits_local_m = _prompt_logits_nograd(model, model.prompt_learner_m, local, ssq, mask, g_unit, b, l,
                                                             logit_scale, spatial, use_evidence)
    return logits, logits_local, image_features, text_features, logits_m, logits_local_m
