"""CPU fp32 restatement of the reference hot path (TEST INFRASTRUCTURE — the checker, never the product).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  The product package never does: it fails loudly when its CUDA library is absent.

Parity status: PINNED.  Every function below is checked against outputs of the reference's own
classes executed in the build container (oracle/ref_extract.py → oracle/make_golden.py →
tests/golden/*.npz; tests/test_oracle_vs_golden.py, and live in tests/test_oracle_vs_reference.py
whenever /root/reference is present).  The reference itself ships no tests or golden vectors
(SURVEY §4), so those reference-generated fixtures are the pin.

Written as plain functions over a CLIP `state_dict` (OpenAI key names) instead of nn.Modules so the
algorithm is stated once, independent of module plumbing.  Citations: T = project/my_code/trainers/
Caption_distill_double.py, M = project/my_code/clip/model.py, U = project/my_code/trainers/utils.py,
EV = project/my_code/Dassl.pytorch-master/dassl/evaluation/evaluator.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# visual trunk: CLIP ModifiedResNet (M:130-190), BatchNorm always in eval mode
# --------------------------------------------------------------------------------------------
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


def _bottleneck(sd, p, x, stride):
    """M:40-53.  Stride is realised as AvgPool2d after conv2 and in front of the downsample conv."""
    y = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    y = F.relu(_bn(sd, p + ".bn2", F.conv2d(y, sd[p + ".conv2.weight"], padding=1)))
    if stride > 1:
        y = F.avg_pool2d(y, stride)
    y = _bn(sd, p + ".bn3", F.conv2d(y, sd[p + ".conv3.weight"]))
    if (p + ".downsample.0.weight") in sd:
        idn = F.avg_pool2d(x, stride) if stride > 1 else x
        idn = _bn(sd, p + ".downsample.1", F.conv2d(idn, sd[p + ".downsample.0.weight"]))
    else:
        idn = x
    return F.relu(y + idn)


def rn_trunk(sd, image, layers):
    """T:385-399 `encode_image` (stem M:173-177 + layer1..4).  [B,3,H,W] -> [B,Cv,H/32,W/32]."""
    x = image.float()
    x = F.relu(_bn(sd, "visual.bn1", F.conv2d(x, sd["visual.conv1.weight"], stride=2, padding=1)))
    x = F.relu(_bn(sd, "visual.bn2", F.conv2d(x, sd["visual.conv2.weight"], padding=1)))
    x = F.relu(_bn(sd, "visual.bn3", F.conv2d(x, sd["visual.conv3.weight"], padding=1)))
    x = F.avg_pool2d(x, 2)
    for li, blocks in enumerate(layers, start=1):
        for bi in range(blocks):
            x = _bottleneck(sd, f"visual.layer{li}.{bi}", x, 2 if (li > 1 and bi == 0) else 1)
    return x


def local_features(sd, feat):
    """T:405-411: attnpool's value->output path on every patch, no positional embedding.
    [B,Cv,h,w] -> [P,B,D]."""
    b, c, h, w = feat.shape
    x = feat.reshape(b, c, h * w).permute(2, 0, 1)
    x = F.linear(x, sd["visual.attnpool.v_proj.weight"], sd["visual.attnpool.v_proj.bias"])
    return F.linear(x, sd["visual.attnpool.c_proj.weight"], sd["visual.attnpool.c_proj.bias"])


def attnpool_global(sd, feat, heads):
    """M:89-127 with `if_pos=False` (T:413), keeping only what reaches token 0:
    tokens = [mean; patches]; q from the mean token only; softmax over P+1 keys per head."""
    b, c, h, w = feat.shape
    x = feat.reshape(b, c, h * w).permute(0, 2, 1)                     # [B,P,C]
    tok = torch.cat([x.mean(dim=1, keepdim=True), x], dim=1)           # [B,P+1,C]
    p = "visual.attnpool."
    q = F.linear(tok[:, 0], sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])         # [B,C]
    k = F.linear(tok, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])               # [B,P+1,C]
    v = F.linear(tok, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])
    dh = c // heads
    q = q.reshape(b, heads, 1, dh) * dh ** -0.5
    k = k.reshape(b, -1, heads, dh).permute(0, 2, 1, 3)
    v = v.reshape(b, -1, heads, dh).permute(0, 2, 1, 3)
    a = torch.softmax(q @ k.transpose(-1, -2), dim=-1)                 # [B,H,1,P+1]
    o = (a @ v).reshape(b, c)
    return F.linear(o, sd[p + "c_proj.weight"], sd[p + "c_proj.bias"])  # [B,D]


# --------------------------------------------------------------------------------------------
# ViT visual tower (BASELINE configs 3, 5).  The reference `DenseCLIP` cannot wrap a VisionTransformer
# (T:365-373 need layer4/attnpool; M:271-276 returns the class token only), so the DENSE output is this
# repo's definition mirroring the ModifiedResNet rule "local = value->output path per patch,
# global = the normally pooled token": in the LAST block every patch token's attention output is
# out_proj(v_proj(ln_1(x))) (no q.k mixing) while the class token attends normally; then the block's
# residual + MLP, ln_post and `proj` on every token.  The GLOBAL feature (token 0) is therefore exactly the
# reference `VisionTransformer.forward` (M:259-276) and is pinned against it (tests/golden/vit_*.npz).
# --------------------------------------------------------------------------------------------
def vit_tokens(sd, image, patch, heads):
    """M:259-276 up to (not including) the last residual block.  [B,3,H,W] -> x [B,1+P,W] after ln_pre and
    blocks 0..L-2, plus the index of the last block."""
    x = F.conv2d(image.float(), sd["visual.conv1.weight"], stride=patch)                  # M:261
    b, w = x.shape[0], x.shape[1]
    x = x.reshape(b, w, -1).permute(0, 2, 1)                                               # [B,P,W]
    cls = sd["visual.class_embedding"].expand(b, 1, w)
    x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"]                     # M:264-265
    x = F.layer_norm(x, (w,), sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"], 1e-5)
    n = 0
    while f"visual.transformer.resblocks.{n}.ln_1.weight" in sd:
        n += 1
    for i in range(n - 1):
        x = _res_block(sd, f"visual.transformer.resblocks.{i}", x, heads, None)
    return x, n - 1


def vit_dense(sd, image, patch, heads):
    """-> (g [B,D] = reference VisionTransformer output, local [P,B,D] = this repo's dense patch features)."""
    x, last = vit_tokens(sd, image, patch, heads)
    p = f"visual.transformer.resblocks.{last}"
    b, t, w = x.shape
    dh = w // heads
    y = F.layer_norm(x, (w,), sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"], 1e-5)
    qkv = F.linear(y, sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"])
    q, k, v = qkv.split(w, dim=-1)
    q0 = q[:, :1].reshape(b, 1, heads, dh).transpose(1, 2) * dh ** -0.5                     # class-token query only
    kh = k.reshape(b, t, heads, dh).transpose(1, 2)
    vh = v.reshape(b, t, heads, dh).transpose(1, 2)
    o0 = (torch.softmax(q0 @ kh.transpose(-1, -2), dim=-1) @ vh).transpose(1, 2).reshape(b, 1, w)
    a = torch.cat([o0, v[:, 1:]], dim=1)                                                    # patches: value path only
    x = x + F.linear(a, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"])
    y = F.layer_norm(x, (w,), sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"], 1e-5)
    x = x + F.linear(_quick_gelu(F.linear(y, sd[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"])),
                     sd[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"])
    x = F.layer_norm(x, (w,), sd["visual.ln_post.weight"], sd["visual.ln_post.bias"], 1e-5) @ sd["visual.proj"]
    return x[:, 0], x[:, 1:].permute(1, 0, 2)


# --------------------------------------------------------------------------------------------
# text tower: T:72-101 TextEncoder over M:207-239 Transformer (causal mask M:364-370)
# --------------------------------------------------------------------------------------------
def _quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)          # M:202-204


def _res_block(sd, p, x, heads, mask):
    """M:207-228.  x [N,L,W]."""
    n, l, w = x.shape
    dh = w // heads
    y = F.layer_norm(x, (w,), sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"], 1e-5)
    qkv = F.linear(y, sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"])
    q, k, v = qkv.split(w, dim=-1)
    q = q.reshape(n, l, heads, dh).transpose(1, 2) * dh ** -0.5
    k = k.reshape(n, l, heads, dh).transpose(1, 2)
    v = v.reshape(n, l, heads, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(n, l, w)
    x = x + F.linear(o, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"])
    y = F.layer_norm(x, (w,), sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"], 1e-5)
    y = F.linear(_quick_gelu(F.linear(y, sd[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"])),
                 sd[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"])
    return x + y


def text_encode(sd, prompts, eot_index=None, heads=8, sequence=False, prefix="transformer"):
    """T:82-101.  `prompts` float [N,L,W] embeddings (pos-emb NOT yet added) -> [N,D] at `eot_index`
    (argmax of the token ids, T:100) or the per-token [N,L,D] when `sequence` (T:94-96)."""
    x = prompts + sd["positional_embedding"]
    l = x.shape[1]
    mask = torch.full((l, l), float("-inf")).triu_(1)
    i = 0
    while f"{prefix}.resblocks.{i}.ln_1.weight" in sd:
        x = _res_block(sd, f"{prefix}.resblocks.{i}", x, heads, mask)
        i += 1
    w = x.shape[-1]
    x = F.layer_norm(x, (w,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    if sequence:
        return x @ sd["text_projection"]
    return x[torch.arange(x.shape[0]), eot_index] @ sd["text_projection"]


def text_encode_adapter(sd, prompts, eot_index, heads, w_down, w_up, prefix="transformer"):
    """`AdapterTextEncoder.forward` (Caption_distill_double_adapter.py:99-125): transformer, x + Adapter(x) with
    Adapter(x) = relu(relu(x W_down^T) W_up^T) (TA:304-317), ln_final, EOT row @ text_projection."""
    x = prompts + sd["positional_embedding"]
    l = x.shape[1]
    mask = torch.full((l, l), float("-inf")).triu_(1)
    i = 0
    while f"{prefix}.resblocks.{i}.ln_1.weight" in sd:
        x = _res_block(sd, f"{prefix}.resblocks.{i}", x, heads, mask)
        i += 1
    x = x + torch.relu(torch.relu(x @ w_down.t()) @ w_up.t())
    x = F.layer_norm(x, (x.shape[-1],), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    return x[torch.arange(x.shape[0]), eot_index] @ sd["text_projection"]


def embed_tokens(sd, token_ids):
    """M:326 token embedding gather."""
    return sd["token_embedding.weight"][token_ids]


def assemble_prompts(prefix, ctx, suffix):
    """T:199-242 for CLASS_TOKEN_POSITION == 'end': [SOS | ctx | class tokens, EOS, pad]."""
    if ctx.dim() == 2:
        ctx = ctx.unsqueeze(0).expand(prefix.shape[0], -1, -1)
    return torch.cat([prefix, ctx, suffix], dim=1)


def _unit(x):
    return x / x.norm(dim=-1, keepdim=True)       # no epsilon, as T:441-442 / T:485-488


# --------------------------------------------------------------------------------------------
# dual-prompt head
# --------------------------------------------------------------------------------------------
def retrieval_mix(g_unit, bank, topk=10):
    """T:444-448.  g <- (g + mean of the top-10 most similar bank rows)/2, not re-normalised.
    Returns (mixed g [B,D], top-k scores [B,10])."""
    sim = g_unit @ bank.float().t()
    scores, idx = sim.topk(topk, -1)
    picked = bank[idx.reshape(-1)].reshape(idx.shape[0], topk, -1).mean(1)
    return torch.cat([g_unit[:, None], picked[:, None].to(g_unit.dtype)], 1).mean(1), scores


def aggregate(neg, evidence, spatial_scale, logit_scale):
    """T:458-470 / T:500-514 on [P,B,K] similarity maps.  Returns (logits_local [B,K], neg map as
    returned by the reference — i.e. after the winner-take-all reweighting when evidence is used)."""
    if evidence is not None:
        w = torch.softmax(spatial_scale * neg * (neg.max(-1)[0].unsqueeze(-1) + 1), -1)
        neg = neg * w
        prob = torch.softmax(evidence * spatial_scale, dim=0)
    else:
        prob = torch.softmax(neg * spatial_scale, dim=0)
    return torch.sum(logit_scale * neg * prob, dim=0), neg


def head_test(g, local, t_pos, t_neg, t_evi=None, bank=None, logit_scale=4.0, spatial_scale=50.0):
    """T:441-472.  g [B,D] raw global feature, local [P,B,D] raw local features, t_* [K,D] *unit* text
    features.  -> (logits_, logits_local, logits_neg [P,B,K], local·t_posᵀ [P,B,K], topk scores|None)."""
    g = _unit(g)
    local = _unit(local)
    scores = None
    if bank is not None:
        g, scores = retrieval_mix(g, bank)
    logits_g = logit_scale * g @ t_pos.t()
    neg = local @ t_neg.t()
    evi = local @ t_evi.t() if t_evi is not None else None
    logits_local, neg = aggregate(neg, evi, spatial_scale, logit_scale)
    return logits_g, logits_local, neg, local @ t_pos.t(), scores


def head_train(seq_feats, captions, t_pos, t_neg, t_evi=None, logit_scale=4.0, spatial_scale=50.0):
    """T:474-514.  seq_feats [B,L,D] per-token text-as-image features, captions [B,L] token ids,
    t_* [K,D] raw prompt features (normalised here, T:487-488,503).
    -> (logits_ [B,K], logits_local [B,K], unit seq feats [L,B,D], unit t_pos [K,D])."""
    b = seq_feats.shape[0]
    g = _unit(seq_feats[torch.arange(b), captions.argmax(dim=-1)])
    local = _unit(seq_feats.permute(1, 0, 2))                    # [L,B,D]
    t_pos, t_neg = _unit(t_pos), _unit(t_neg)
    mask = (captions == 0).to(local.dtype) * (-10000.0)          # [B,L]; token id 0 = padding (T:491)
    logits_g = logit_scale * g @ t_pos.t()
    neg = local @ t_neg.t() + mask.t()[:, :, None]               # [L,B,K]
    evi = None
    if t_evi is not None:
        evi = local @ _unit(t_evi).t() + mask.t()[:, :, None]
    logits_local, _ = aggregate(neg, evi, spatial_scale, logit_scale)
    return logits_g, logits_local, local, t_pos


# --------------------------------------------------------------------------------------------
# whole-model drivers (what DenseCLIP.forward does, T:401-545), over a state_dict
# --------------------------------------------------------------------------------------------
def prompt_features(sd, prompt_learner_state, token_ids, heads=8, use_evidence=False):
    """Encode the 80 pos / neg (/ evidence) prompts (T:428-438 / T:480-483,502).  Raw (un-normalised)."""
    pl = prompt_learner_state
    eot = token_ids.argmax(dim=-1)
    out = []
    for key in ("ctx", "ctx_double") + (("ctx_evidence",) if use_evidence else ()):
        emb = assemble_prompts(pl["token_prefix"], pl[key], pl["token_suffix"])
        out.append(text_encode(sd, emb, eot, heads))
    return out


def prompt_learner_state(sd, token_ids, n_ctx, ctx, ctx_double, ctx_evidence):
    """T:176-191: frozen SOS prefix / class+EOS suffix embeddings around the learnable contexts."""
    emb = embed_tokens(sd, token_ids)
    return {"token_prefix": emb[:, :1], "token_suffix": emb[:, 1 + n_ctx:],
            "ctx": ctx, "ctx_double": ctx_double, "ctx_evidence": ctx_evidence}


def dense_clip_test(sd, arch, image, pl_state, token_ids, use_evidence=False, bank=None,
                    logit_scale=4.0, spatial_scale=50.0):
    """DenseCLIP.forward(image, if_test=True) (T:402-472) for a ModifiedResNet CLIP."""
    feat = rn_trunk(sd, image, arch.vision_layers)
    local = local_features(sd, feat)
    g = attnpool_global(sd, feat, arch.vision_width * 32 // 64)
    feats = [_unit(t) for t in prompt_features(sd, pl_state, token_ids, arch.transformer_heads, use_evidence)]
    t_evi = feats[2] if use_evidence else None
    return head_test(g, local, feats[0], feats[1], t_evi, bank, logit_scale, spatial_scale)


def dense_clip_test_vit(sd, arch, image, pl_state, token_ids, use_evidence=False, bank=None,
                        logit_scale=4.0, spatial_scale=50.0):
    """The same head (T:441-472) on the ViT tower's dense tokens (repo-defined, see `vit_dense`)."""
    g, local = vit_dense(sd, image, arch.vision_patch_size, arch.vision_width // 64)
    feats = [_unit(t) for t in prompt_features(sd, pl_state, token_ids, arch.transformer_heads, use_evidence)]
    t_evi = feats[2] if use_evidence else None
    return head_test(g, local, feats[0], feats[1], t_evi, bank, logit_scale, spatial_scale)


def dense_clip_train(sd, arch, captions, pl_state, token_ids, use_evidence=False,
                     logit_scale=4.0, spatial_scale=50.0):
    """DenseCLIP.forward(None, captions) (T:473-545, ema off): 4 leading outputs."""
    seq = text_encode(sd, embed_tokens(sd, captions), None, arch.transformer_heads, sequence=True)
    feats = prompt_features(sd, pl_state, token_ids, arch.transformer_heads, use_evidence)
    t_evi = feats[2] if use_evidence else None
    return head_train(seq, captions, feats[0], feats[1], t_evi, logit_scale, spatial_scale)


def adapter_dense_clip_test(sd, arch, image, pl_state, token_ids, w_down, w_up, logit_scale=4.0, spatial_scale=50.0):
    """`AdapterDenseCLIP.forward(image, if_test=True)` (Caption_distill_double_adapter.py:365-412): no evidence, no retrieval."""
    feat = rn_trunk(sd, image, arch.vision_layers)
    local = local_features(sd, feat)
    g = attnpool_global(sd, feat, arch.vision_width * 32 // 64)
    eot = token_ids.argmax(dim=-1)
    tf = [_unit(text_encode_adapter(sd, assemble_prompts(pl_state["token_prefix"], pl_state[k], pl_state["token_suffix"]), eot,
                                    arch.transformer_heads, w_down, w_up)) for k in ("ctx", "ctx_double")]
    return head_test(g, local, tf[0], tf[1], None, None, logit_scale, spatial_scale)[:4]


def adapter_dense_clip_train(sd, arch, captions, pl_state, token_ids, w_down, w_up, logit_scale=4.0, spatial_scale=50.0):
    """`AdapterDenseCLIP.forward(None, captions)` (Caption_distill_double_adapter.py:413-457): captions through the plain text
    encoder, prompts through the adapter encoder."""
    seq = text_encode(sd, embed_tokens(sd, captions), None, arch.transformer_heads, sequence=True)
    eot = token_ids.argmax(dim=-1)
    tf = [text_encode_adapter(sd, assemble_prompts(pl_state["token_prefix"], pl_state[k], pl_state["token_suffix"]), eot,
                              arch.transformer_heads, w_down, w_up) for k in ("ctx", "ctx_double")]
    return head_train(seq, captions, tf[0], tf[1], None, logit_scale, spatial_scale)


# --------------------------------------------------------------------------------------------
# losses (U:85-190)
# --------------------------------------------------------------------------------------------
def ranking_loss(y_pred, y_true, scale=2.0, margin=1.0):
    """U:85-93 without the in-place `y_pred *= scale_` side effect:
    mean_b sum_{i,j} relu(margin - s*y_j + s*y_i) * t_j * (1 - t_i)."""
    y = y_pred * scale
    t = y_true.float()
    d = (margin - y[:, None, :] + y[:, :, None]).clamp_min(0)
    return (d * t[:, None, :] * (1 - t[:, :, None])).sum((-1, -2)).mean()


def asymmetric_loss(x, y, gamma_neg, gamma_pos, clip, eps, thresh_pos, thresh_neg, partial):
    """U:136-173.  The focal weight (1-p_t)^gamma is computed with grad disabled (U:162-170)."""
    p = torch.sigmoid(x)
    pneg = 1 - p
    if clip is not None and clip > 0:
        pneg = (pneg + clip).clamp(max=1)
    ypos = (y > thresh_pos).float()
    yneg = (y < thresh_neg).float()
    loss = ypos * torch.log(p.clamp(min=eps)) + yneg * torch.log(pneg.clamp(min=eps))
    if gamma_neg > 0 or gamma_pos > 0:
        with torch.no_grad():
            pt = p * ypos + pneg * yneg
            wgt = torch.pow(1 - pt, gamma_pos * ypos + gamma_neg * yneg)
        loss = loss * wgt
    return -loss.sum() / x.shape[0] if partial else -loss.mean()


def asl_loss(x, y):
    """U:184-190 `ASL_loss`."""
    return asymmetric_loss(x, y, 2, 1, 0.05, 1e-8, 0.9, 0.9, False)


def dualcoop_loss(x, y):
    """U:175-181 `dualcoop_loss` (partial labels: targets in {-1,0,1})."""
    return asymmetric_loss(x, y, 2, 1, 0.05, 1e-8, 0.9, -0.9, True)


def ranking_loss_with_cooccurrence(y_pred, y_true, cooccurrence, scale=2.0, margin=1.0):
    """U:95-110: the hinge of pair (i, j) weighted by log(1 / (p_ij + 1e-6)), zero diagonal, rows divided by their mean."""
    y = y_pred * scale
    t = y_true.float()
    w = (1 / (cooccurrence + 1e-6)).log()
    w = w * (1 - torch.eye(w.shape[0], w.shape[1]))
    w = w / w.mean(-1)[:, None]
    d = (margin - y[:, None, :] + y[:, :, None]).clamp_min(0) * w
    return (d * t[:, None, :] * (1 - t[:, :, None])).sum((-1, -2)).mean()


def resample_loss(x, y, class_freq, neg_class_freq, reweight=True, map_alpha=10.0, map_beta=0.2, map_gamma=0.1, logit_reg=None,
                  focal=False, focal_gamma=2.0, balance_param=2.0, loss_weight=1.0):
    """`ResampleLoss.forward` (dbl.py:351-383) for use_sigmoid=True, partial=False, reweight_func None | 'rebalance':
    every `binary_cross_entropy` call reduces with 'mean' (dbl.py:61-63), so the focal factor multiplies two scalars."""
    logit_reg = logit_reg or {}
    y = y.float()
    freq_inv = 1.0 / class_freq
    w = None
    if reweight:
        repeat_rate = (y * freq_inv).sum(1, keepdim=True)
        w = torch.sigmoid(map_beta * (freq_inv[None] / repeat_rate - map_gamma)) + map_alpha
    neg_scale = logit_reg.get("neg_scale", 1.0)
    if "init_bias" in logit_reg:
        train_num = class_freq[0] + neg_class_freq[0]
        x = x + (-torch.log(train_num / class_freq - 1) * logit_reg["init_bias"] / neg_scale)
    if "neg_scale" in logit_reg:
        x = x * (1 - y) * neg_scale + x * y
        w = w / neg_scale * (1 - y) + w * y
    bce = torch.nn.functional.binary_cross_entropy_with_logits
    if focal:
        pt = torch.exp(-bce(x, y, None, reduction="mean"))
        return loss_weight * balance_param * (1 - pt) ** focal_gamma * bce(x, y, w, reduction="mean")
    return loss_weight * bce(x, y, w, reduction="mean")


def kl_softmax(x, x_target):
    """nn.KLDivLoss(reduction="batchmean")(log_softmax(x), softmax(x_target)) (T:796, T:809-811): sum_b,k q (log q - log p) / B."""
    logp = torch.log_softmax(x, dim=-1)
    q = torch.softmax(x_target, dim=-1)
    return torch.xlogy(q, q).sum() / x.shape[0] - (q * logp).sum() / x.shape[0]


def ema_loss(out, out_m, out_local, out_local_m):
    """T:809-811."""
    return kl_softmax(out, out_m) + kl_softmax(out_local, out_local_m) * 10000


def momentum_update(live, twin, momentum=0.995):
    """T:554-559 over dicts of tensors: twin <- twin * m + live * (1 - m)."""
    return {k: twin[k] * momentum + live[k] * (1.0 - momentum) for k in twin}


def dense_clip_train_ema(sd, arch, captions, pl_state, pl_state_m, token_ids, use_evidence=False, logit_scale=4.0,
                         spatial_scale=50.0, momentum=0.995):
    """DenseCLIP.forward(None, captions) with TRAIN.ema (T:516-541): the twin is updated FIRST (T:518), then encodes its own
    prompts without gradient.  -> (logits_, logits_local, feats, t_pos, logits_m_, logits_local_m, updated twin state)."""
    out = dense_clip_train(sd, arch, captions, pl_state, token_ids, use_evidence, logit_scale, spatial_scale)
    with torch.no_grad():
        keys = ("ctx", "ctx_double", "ctx_evidence")
        new_m = dict(pl_state_m)
        new_m.update(momentum_update({k: pl_state[k].detach() for k in keys}, {k: pl_state_m[k] for k in keys}, momentum))
        out_m = dense_clip_train(sd, arch, captions, new_m, token_ids, use_evidence, logit_scale, spatial_scale)
    return out + (out_m[0], out_m[1], new_m)


# --------------------------------------------------------------------------------------------
# metric (EV:137-175)
# --------------------------------------------------------------------------------------------
def mean_average_precision(targets: np.ndarray, scores: np.ndarray) -> float:
    """EV:137-175: per-class AP over examples sorted by descending score (eps 1e-8), x100."""
    if scores.size == 0:
        return 0.0
    aps = []
    for k in range(scores.shape[1]):
        order = np.argsort(scores[:, k])[::-1]
        hit = targets[order, k] == 1
        cum = np.cumsum(hit)
        prec = np.where(hit, cum, 0) / np.arange(1, len(order) + 1)
        aps.append(prec.sum() / (cum[-1] + 1e-8))
    return float(100.0 * np.mean(aps))


def aggregate_blocks(output, output_blocks, threshold=0.3, weight=1.4):
    """T:655-662: per class, the max over the sliding windows if it exceeds the threshold, else the min."""
    alpha = output_blocks.max(dim=1)[0]
    beta = output_blocks.min(dim=1)[0]
    gamma = (alpha > threshold).int()
    return weight * (gamma * alpha + (1 - gamma) * beta) + output


def _max_min_threshold(data, threshold):
    alpha, beta = data.max(dim=1)[0], data.min(dim=1)[0]
    gamma = (alpha > threshold).int()
    return gamma * alpha + (1 - gamma) * beta


def fuse(data, sims_scores, threshold=0.2):
    """gen_final_ans.py:18-36: windows re-weighted by 1 + mean similarity, then by 1 + var_k (unbiased)."""
    data = (1 + sims_scores.mean(-1, keepdim=True)) * data
    data = (1 + torch.var(data, dim=2).unsqueeze(-1)) * data
    return _max_min_threshold(data, threshold)


def fuse6(data, sims_scores, threshold=0.2):
    """gen_final_ans.py:38-71."""
    v0 = 1 + torch.var(data, dim=2).unsqueeze(-1)
    d_sim = (1 + sims_scores.mean(-1, keepdim=True)) * data
    v1 = 1 + torch.var(d_sim, dim=2).unsqueeze(-1)
    return _max_min_threshold(v0 * v1 * d_sim, threshold)


def cooccurrence_adjust(pred, adj, nums, weight=0.5):
    """T:614-618,632-636: pred + w * pred @ rownorm(adj / nums)."""
    p = torch.as_tensor(adj / nums[:, None], dtype=torch.float32)
    p = p / p.sum(-1)[:, None]
    return pred + weight * pred @ p
