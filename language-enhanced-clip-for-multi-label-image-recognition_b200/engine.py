"""B200 execution engines for the two frozen CLIP towers.

`VisualRN`  — CLIP ModifiedResNet trunk + dense/global pooling (reference: T:385-413, M:10-190).
`TextTower` — CLIP text transformer (reference: T:72-101, M:207-239).

Data layout in HBM: activations are NHWC bf16, i.e. row-major [B*H*W, C] matrices, so every 1x1
convolution / linear layer is one `lecb_gemm_bf16` launch and every 3x3 convolution one
`lecb_conv3x3_bf16` launch (TMA im2col) with eval-BatchNorm folded into the bf16 weights and an fp32
bias, ReLU / residual-add fused in the epilogue.  Weights are packed once at construction.
"""
from __future__ import annotations

import os

import torch

from . import ops


def _fold_bn(sd, conv_key, bn_prefix, eps=1e-5):
    """eval BatchNorm folded into the preceding conv (BN is always in eval mode: SURVEY §3.3)."""
    w = sd[conv_key].float()
    scale = sd[bn_prefix + ".weight"].float() / torch.sqrt(sd[bn_prefix + ".running_var"].float() + eps)
    bias = sd[bn_prefix + ".bias"].float() - sd[bn_prefix + ".running_mean"].float() * scale
    return w * scale.view(-1, 1, 1, 1), bias.contiguous()


def _w1x1(w):
    return w.reshape(w.shape[0], w.shape[1]).to(torch.bfloat16).contiguous()


def _w3x3(w):
    return w.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()          # [Cout,3,3,Cin]


def _pack_pixel_pairs(w, bias):
    """Re-express a 3x3 / pad-1 conv over NHWC [B,H,W,Ci] as the same kind of conv over the *pixel-pair*
    view [B,H,W/2,2*Ci] -> [B,H,W/2,2*Co] (identical memory).  Used for the Ci = 32 stem convolutions:
    64-byte im2col rows and K steps of 32 leave the tensor pipe and the TMA unit badly under-fed; the pair
    view gives 128-byte rows, BK = 64 and twice the MMA N at the price of 50 % structural zeros in the
    (tiny) weight matrix.  w fp32 [Co,3,3,Ci] (ky,kx,ci) -> [2*Co,3,3,2*Ci]; output pixel 2j+o reads input
    pixel 2j+o+kx-1 = pair j + floor((o+kx-1)/2), element (o+kx-1) mod 2."""
    co, _, _, ci = w.shape
    out = w.new_zeros((2, co, 3, 3, 2, ci))
    for o in range(2):
        for kx in range(3):
            d = o + kx - 1
            pt, ip = d // 2 + 1, d % 2
            out[o, :, :, pt, ip, :] = w[:, :, kx, :]
    return out.reshape(2 * co, 3, 3, 2 * ci), torch.cat([bias, bias]).contiguous()


class VisualRN:
    def __init__(self, sd, layers, width, heads, embed_dim, device):
        sd = {k: v.detach().to(device) for k, v in sd.items() if k.startswith("visual.")}
        self.layers, self.width, self.heads, self.embed_dim, self.device = tuple(layers), width, heads, embed_dim, device
        w, b = _fold_bn(sd, "visual.conv1.weight", "visual.bn1")
        self.stem1 = (w.permute(1, 2, 3, 0).reshape(27, -1).contiguous(), b)          # fp32 [27,Cout]
        # Ci = 32 stem convs run directly (halo-tile conv with 64-byte rows, gemm_tcgen05.cu); narrower test towers
        # (Ci = 16) use the pixel-pair view to reach the kernel's Cin % 32 == 0 granularity
        self.stem_pairs = ((width // 2) % 32 != 0) or bool(os.environ.get("LECB_STEM_PAIRS"))
        for name, conv, bn in (("stem2", "visual.conv2.weight", "visual.bn2"), ("stem3", "visual.conv3.weight", "visual.bn3")):
            w, b = _fold_bn(sd, conv, bn)
            w = w.permute(0, 2, 3, 1).contiguous()                                    # [Co,3,3,Ci] fp32
            if self.stem_pairs:
                w, b = _pack_pixel_pairs(w, b)
            setattr(self, name, (w.to(torch.bfloat16).contiguous(), b))
        self.blocks = []
        for li, nblk in enumerate(self.layers, start=1):
            for bi in range(nblk):
                p = f"visual.layer{li}.{bi}"
                blk = {"stride": 2 if (li > 1 and bi == 0) else 1}
                w, b = _fold_bn(sd, p + ".conv1.weight", p + ".bn1")
                blk["c1"] = (_w1x1(w), b)
                w, b = _fold_bn(sd, p + ".conv2.weight", p + ".bn2")
                blk["c2"] = (_w3x3(w), b)
                w, b = _fold_bn(sd, p + ".conv3.weight", p + ".bn3")
                blk["c3"] = (_w1x1(w), b)
                if (p + ".downsample.0.weight") in sd:
                    w, b = _fold_bn(sd, p + ".downsample.0.weight", p + ".downsample.1")
                    blk["ds"] = (_w1x1(w), b)
                    # M:46-52 as ONE K-concatenated GEMM: relu([y | idn] . [W3 | Wd]^T + b3 + bd) — the shortcut tensor is never
                    # written (ops.gemm_dual); needs 64-channel granularity on both operands, else the two-GEMM form below
                    w3, b3 = blk["c3"]
                    if w3.shape[1] % 64 == 0 and blk["ds"][0].shape[1] % 64 == 0 and not os.environ.get("LECB_NO_DUAL"):
                        blk["c3ds"] = (torch.cat([w3, blk["ds"][0]], 1).contiguous(), (b3 + b).contiguous())
                self.blocks.append(blk)
        ap = "visual.attnpool."
        self.proj = {n: (sd[ap + n + ".weight"].to(torch.bfloat16).contiguous(), sd[ap + n + ".bias"].float().contiguous())
                     for n in ("q_proj", "k_proj", "v_proj", "c_proj")}

    # ---- T:385-399 encode_image ----
    def trunk(self, image, mean=None, std=None):
        """image fp32 NCHW [B,3,H,W] (normalised, the reference's tensor) or uint8 NHWC [B,H,W,3] (raw pixels, normalised
        with mean / std inside the stem kernel) -> layer4 features, NHWC bf16 [B,H/32,W/32,Cv]."""
        if image.dtype == torch.uint8:
            x = ops.stem_conv1_u8(image.contiguous(), *self.stem1, mean=mean or ops.CLIP_PIXEL_MEAN, std=std or ops.CLIP_PIXEL_STD)
        else:
            x = ops.stem_conv1(image.contiguous(), *self.stem1)
        b, h, w, c = x.shape
        if self.stem_pairs:
            x = x.view(b, h, w // 2, 2 * c)
        from . import prof
        prof.ALGO_FLOP_SCALE = 0.5 if self.stem_pairs else 1.0     # half of the packed MACs are structural zeros
        x = ops.conv3x3(x, *self.stem2)
        if self.stem_pairs:
            x = ops.conv3x3(x, *self.stem3)
            prof.ALGO_FLOP_SCALE = 1.0
            x = ops.avgpool2x2(x.view(b, h, w, -1))
        else:
            x = ops.conv3x3(x, *self.stem3, pool=True)        # M:147 avgpool fused into the conv epilogue when it can be
        for blk in self.blocks:
            x = self._bottleneck(x, blk)
        return x

    @staticmethod
    def _bottleneck(x, blk):
        b, h, w, c = x.shape
        y = ops.gemm(x.view(-1, c), *blk["c1"], relu=True)
        y = ops.conv3x3(y.view(b, h, w, -1), *blk["c2"])
        idn = x
        if blk["stride"] > 1:
            y = ops.avgpool2x2(y)
            idn = ops.avgpool2x2(x)
            h, w = h // 2, w // 2
        if "c3ds" in blk:
            out = ops.gemm_dual(y.view(b * h * w, -1), idn.view(b * h * w, -1), *blk["c3ds"], relu=True)
            return out.view(b, h, w, -1)
        if "ds" in blk:
            idn = ops.gemm(idn.view(-1, c), *blk["ds"])
        out = ops.gemm(y.view(b * h * w, -1), *blk["c3"], residual=idn.view(b * h * w, -1), relu=True)
        return out.view(b, h, w, -1)

    # ---- T:405-413: per-patch value->output path + single-query attention pool ----
    def pooled(self, feat):
        """feat NHWC bf16 [B,h,w,Cv] -> (local bf16 [B*P,D] un-normalised, row_sumsq fp32 [B*P], g fp32 [B,D])."""
        b, h, w, c = feat.shape
        p = h * w
        x2 = feat.view(b * p, c)
        v = ops.gemm(x2, *self.proj["v_proj"])                         # shared by local path and attnpool values
        ssq = torch.zeros((b * p,), device=feat.device, dtype=torch.float32)
        local = ops.gemm(v, *self.proj["c_proj"], row_sumsq=ssq)
        k = ops.gemm(x2, *self.proj["k_proj"])
        mean_tok = ops.token_mean(feat.view(b, p, c))
        q = ops.gemm(mean_tok, *self.proj["q_proj"], out_f32=True)
        o = ops.attnpool_query0(q, k, v, b, p, self.heads)
        g = ops.gemm(o, *self.proj["c_proj"], out_f32=True)
        return local, ssq, g


def _tower_blocks(g, prefix, layers):
    out = []
    for i in range(layers):
        p = f"{prefix}.resblocks.{i}"
        out.append({
            "ln1": (g(p + ".ln_1.weight").float().contiguous(), g(p + ".ln_1.bias").float().contiguous()),
            "ln2": (g(p + ".ln_2.weight").float().contiguous(), g(p + ".ln_2.bias").float().contiguous()),
            "qkv": (g(p + ".attn.in_proj_weight").to(torch.bfloat16).contiguous(), g(p + ".attn.in_proj_bias").float().contiguous()),
            "out": (g(p + ".attn.out_proj.weight").to(torch.bfloat16).contiguous(), g(p + ".attn.out_proj.bias").float().contiguous()),
            "fc": (g(p + ".mlp.c_fc.weight").to(torch.bfloat16).contiguous(), g(p + ".mlp.c_fc.bias").float().contiguous()),
            "proj": (g(p + ".mlp.c_proj.weight").to(torch.bfloat16).contiguous(), g(p + ".mlp.c_proj.bias").float().contiguous()),
        })
    return out


class VisualViT:
    """CLIP VisionTransformer tower (M:240-276) with this repo's dense last block (oracle `vit_dense`):
    patch tokens of the last block take out_proj(v_proj(ln_1 x)) as their attention output (value path only,
    the ViT analogue of T:405-411), the class token attends normally (tcgen05 attention kernel, one query row).
    Residual stream fp32 [B*T, W]; GEMM operands bf16; token row b*T + 0 is the class token."""

    def __init__(self, sd, patch, width, layers, heads, embed_dim, device):
        g = lambda k: sd[k].detach().to(device)
        self.patch, self.width, self.nlayers, self.heads, self.embed_dim, self.device = patch, width, layers, heads, embed_dim, device
        k = 3 * patch * patch
        self.kpad = (k + 63) // 64 * 64
        w = torch.zeros((width, self.kpad), device=device, dtype=torch.float32)
        w[:, :k] = g("visual.conv1.weight").float().reshape(width, k)
        self.patch_w = w.to(torch.bfloat16).contiguous()
        self.cls = g("visual.class_embedding").float().contiguous()
        self.pos = g("visual.positional_embedding").float().contiguous()
        self.ln_pre = (g("visual.ln_pre.weight").float().contiguous(), g("visual.ln_pre.bias").float().contiguous())
        self.ln_post = (g("visual.ln_post.weight").float().contiguous(), g("visual.ln_post.bias").float().contiguous())
        self.proj_t = g("visual.proj").t().to(torch.bfloat16).contiguous()           # [D, W] K-major
        self.blocks = _tower_blocks(g, "visual.transformer", layers)

    def _mlp(self, x, blk):
        h, _, _, _ = ops.layernorm(x, *blk["ln2"])
        u = ops.gemm(h, *blk["fc"], quick_gelu=True)
        return ops.gemm_f32res(u, *blk["proj"], x)

    def tokens(self, image):
        """image fp32 NCHW [B,3,H,W] -> (feat bf16 [B*T, D] = ln_post(x) @ proj for every token, ssq fp32 [B*T], T)."""
        b, _, hh, ww = image.shape
        t = (hh // self.patch) * (ww // self.patch) + 1
        assert t == self.pos.shape[0], f"image {hh}x{ww} does not match the positional embedding ({self.pos.shape[0]} tokens)"
        w = self.width
        emb = ops.gemm(ops.patchify(image.contiguous(), self.patch, self.kpad), self.patch_w)       # M:261
        x = ops.vit_embed_ln(emb, self.cls, self.pos, *self.ln_pre, b, t)                           # M:262-266
        for i, blk in enumerate(self.blocks):
            h, _, _, _ = ops.layernorm(x, *blk["ln1"])
            qkv = ops.gemm(h, *blk["qkv"])
            if i + 1 < self.nlayers:
                a = ops.attn_fwd(qkv, b, t, w, self.heads)
            else:                                              # dense last block
                a = ops.copy_cols(qkv, 2 * w, w)               # every token: its own value vector
                ops.attn_fwd(qkv, b, t, w, self.heads, q_rows=1, out=a)      # class token: real attention
            x = ops.gemm_f32res(a, *blk["out"], x)
            x = self._mlp(x, blk)
        h, _, _, _ = ops.layernorm(x, *self.ln_post)
        ssq = torch.zeros((b * t,), device=image.device, dtype=torch.float32)
        feat = ops.gemm(h, self.proj_t, row_sumsq=ssq)
        return feat, ssq, t


class TextTower:
    def __init__(self, sd, width, heads, layers, embed_dim, device):
        g = lambda k: sd[k].detach().to(device)
        self.width, self.heads, self.nlayers, self.embed_dim, self.device = width, heads, layers, embed_dim, device
        self.pos = g("positional_embedding").float().contiguous()
        self.tok = g("token_embedding.weight").float().contiguous()
        self.blocks = []
        for i in range(layers):
            p = f"transformer.resblocks.{i}"
            self.blocks.append({
                "ln1": (g(p + ".ln_1.weight").float().contiguous(), g(p + ".ln_1.bias").float().contiguous()),
                "ln2": (g(p + ".ln_2.weight").float().contiguous(), g(p + ".ln_2.bias").float().contiguous()),
                "qkv": (g(p + ".attn.in_proj_weight").to(torch.bfloat16).contiguous(), g(p + ".attn.in_proj_bias").float().contiguous()),
                "out": (g(p + ".attn.out_proj.weight").to(torch.bfloat16).contiguous(), g(p + ".attn.out_proj.bias").float().contiguous()),
                "fc": (g(p + ".mlp.c_fc.weight").to(torch.bfloat16).contiguous(), g(p + ".mlp.c_fc.bias").float().contiguous()),
                "proj": (g(p + ".mlp.c_proj.weight").to(torch.bfloat16).contiguous(), g(p + ".mlp.c_proj.bias").float().contiguous()),
            })
        self.ln_final = (g("ln_final.weight").float().contiguous(), g("ln_final.bias").float().contiguous())
        self.text_proj_t = g("text_projection").t().to(torch.bfloat16).contiguous()      # [D, W] K-major
        self.adapter = None

    def set_adapter(self, w_down, w_up):
        """Frozen residual adapter between the transformer and ln_final (Caption_distill_double_adapter.py:304-317, :109):
        x <- x + relu(relu(x W_down^T) W_up^T).  Only the EOT row of a prompt reaches the output (ln_final and the projection
        are per-row), so the adapter runs on those N rows.  w_down [W/4, W], w_up [W, W/4] (nn.Linear weights, no bias)."""
        wd, wu = w_down.detach().to(self.device), w_up.detach().to(self.device)
        self.adapter = {"down": wd.to(torch.bfloat16).contiguous(), "up": wu.to(torch.bfloat16).contiguous(),
                        "down_t": wd.t().to(torch.bfloat16).contiguous(), "up_t": wu.t().to(torch.bfloat16).contiguous()}

    def _adapter_fwd(self, xr):
        """xr fp32 [N,W] -> (xr + adapter(xr), a1 bf16 [N,W/4] post-ReLU, z2 fp32 [N,W] pre-ReLU)."""
        a1 = ops.gemm(xr.to(torch.bfloat16), self.adapter["down"], relu=True)
        z2 = ops.gemm(a1, self.adapter["up"], out_f32=True)
        return ops.residual_relu_fwd(xr, z2), a1, z2

    def _adapter_bwd(self, d_out, a1, z2):
        """d_out fp32 [N,W] = dL/d(xr + adapter(xr)) -> dL/dxr (identity path + data gradient through the frozen adapter)."""
        dz2 = ops.relu_bwd(d_out, z2)                                        # [N,W] bf16
        da1 = ops.gemm(dz2, self.adapter["up_t"], out_f32=True)             # [N,W/4] = dz2 @ W_up
        dz1 = ops.relu_bwd(da1, a1)
        return ops.gemm_f32res(dz1, self.adapter["down_t"], None, d_out.contiguous())      # dz1 @ W_down + d_out

    # ------------------------------------------------------------------ prompt-tuning (fwd with saves + bwd)
    def _dgrad_weights(self):
        """Transposed copies of the frozen weights for the data-gradient GEMMs (built on first use)."""
        if getattr(self, "_wt", None) is None:
            self._wt = [{k: blk[k][0].t().contiguous() for k in ("qkv", "out", "fc", "proj")} for blk in self.blocks]
            self._text_proj = self.text_proj_t.t().contiguous()              # [W, D]: W-operand of d_rows = dT @ P^T
        return self._wt

    def forward_train(self, x, eot_index, rider=None):
        """Like forward(x, eot_index) but keeps what the backward needs.  x fp32 [N,L,W] -> (T_raw fp32 [N,D], saved).

        `rider` = fp32 [Nc,Lc,W]: a second, gradient-free batch of sequences (the caption branch of the prompt-tuning step,
        T:474-477) carried through the SAME per-layer launches — LayerNorm, the four GEMMs and QuickGELU are row-wise, so
        its rows are appended to the prompt rows and only the attention runs once per group (different N and L).  Halves the
        launches of a step that is bound by them.  Returns (T_raw, saved, rider rows before ln_final fp32 [Nc*Lc,W])."""
        n, l, w = x.shape
        x = x.reshape(n * l, w)
        rp = n * l
        if rider is not None:
            nc, lc, _ = rider.shape
            x = torch.cat([x, rider.reshape(nc * lc, w)], 0)
        x = x.contiguous()
        saved = {"n": n, "l": l, "w": w, "eot": eot_index, "layers": []}
        for blk in self.blocks:
            h, _, m1, r1 = ops.layernorm(x, *blk["ln1"], save_stats=True)
            qkv = ops.gemm(h, *blk["qkv"])
            if rider is None:
                a = ops.causal_attn(qkv, n, l, w, self.heads)  # bwd (lecb_attn_causal_bwd) rounds P to bf16 the same way
            else:
                a = torch.empty((x.shape[0], w), device=x.device, dtype=torch.bfloat16)
                ops.attn_fwd(qkv[:rp], n, l, w, self.heads, causal=True, out=a[:rp])
                ops.attn_fwd(qkv[rp:], nc, lc, w, self.heads, causal=True, out=a[rp:])
            x1 = ops.gemm_f32res(a, *blk["out"], x)
            h, _, m2, r2 = ops.layernorm(x1, *blk["ln2"], save_stats=True)
            v = ops.gemm(h, *blk["fc"])                         # pre-activation kept for the QuickGELU backward
            u = ops.quick_gelu_fwd(v)
            x2 = ops.gemm_f32res(u, *blk["proj"], x1)
            saved["layers"].append((x[:rp], m1[:rp], r1[:rp], qkv[:rp], x1[:rp], m2[:rp], r2[:rp], v[:rp]))
            x = x2
        tail = None
        if rider is not None:
            tail, x = x[rp:], x[:rp]
        if self.adapter is not None:
            xr = x.view(n, l, w)[torch.arange(n, device=x.device), eot_index].contiguous()
            xr2, a1, z2 = self._adapter_fwd(xr)
            h, _, mf, rf = ops.layernorm(xr2, *self.ln_final, save_stats=True)
            saved["final"] = (xr2, mf, rf)
            saved["adapter"] = (a1, z2)
            t_raw = ops.gemm(h, self.text_proj_t, out_f32=True)
        else:
            h, _, mf, rf = ops.layernorm(x, *self.ln_final, save_stats=True)
            saved["final"] = (x, mf, rf)
            rows = h.view(n, l, w)[torch.arange(n, device=h.device), eot_index].contiguous()
            t_raw = ops.gemm(rows, self.text_proj_t, out_f32=True)
        return (t_raw, saved) if rider is None else (t_raw, saved, tail)

    def backward(self, saved, d_out):
        """d_out fp32 [N,D] = dL/dT_raw  ->  dL/dx fp32 [N,L,W] (x = prompt embeddings + positional embedding)."""
        wt = self._dgrad_weights()
        n, l, w = saved["n"], saved["l"], saved["w"]
        d_rows = ops.gemm(d_out.to(torch.bfloat16).contiguous(), self._text_proj, out_f32=True)      # [N,W]
        xf, mf, rf = saved["final"]
        if "adapter" in saved:
            d_xr2, _ = ops.layernorm_bwd(d_rows, xf, self.ln_final[0], mf, rf, want_bf16=False)       # N rows
            d_xr = self._adapter_bwd(d_xr2, *saved["adapter"])
            dx = torch.zeros((n, l, w), device=d_out.device, dtype=torch.float32)
            dx[torch.arange(n, device=d_out.device), saved["eot"]] = d_xr                             # only EOT rows see the loss
            dx = dx.view(n * l, w)
            dxb = dx.to(torch.bfloat16)
        else:
            dh = torch.zeros((n, l, w), device=d_out.device, dtype=torch.float32)
            dh[torch.arange(n, device=d_out.device), saved["eot"]] = d_rows                           # only EOT rows see the loss
            dx, dxb = ops.layernorm_bwd(dh.view(n * l, w), xf, self.ln_final[0], mf, rf)
        for blk, t, (x0, m1, r1, qkv, x1, m2, r2, v) in zip(reversed(self.blocks), reversed(wt), reversed(saved["layers"])):
            dv = ops.gemm_mul_quick_gelu_grad(dxb, t["proj"], v)          # [M,4W] bf16: (dx @ W_proj) * QuickGELU'(v)
            dh2 = ops.gemm(dv, t["fc"], out_f32=True)
            dx1, dx1b = ops.layernorm_bwd(dh2, x1, blk["ln2"][0], m2, r2, dx_in=dx)
            da = ops.gemm(dx1b, t["out"])
            dqkv = ops.causal_attn_bwd(qkv, da, n, l, w, self.heads)
            dh1 = ops.gemm(dqkv, t["qkv"], out_f32=True)
            dx, dxb = ops.layernorm_bwd(dh1, x0, blk["ln1"][0], m1, r1, dx_in=dx1)
        return dx.view(n, l, w)

    def forward(self, x, eot_index=None, sequence=False):
        """T:82-101.  x fp32 [N,L,W] = embeddings + positional embedding.  -> fp32 [N,D] at `eot_index`
        (int64 [N]) or fp32 [N,L,D] when `sequence`."""
        n, l, w = x.shape
        x = x.reshape(n * l, w).contiguous()
        for blk in self.blocks:
            h, _, _, _ = ops.layernorm(x, *blk["ln1"])
            qkv = ops.gemm(h, *blk["qkv"])
            a = ops.causal_attn(qkv, n, l, w, self.heads)
            x = ops.gemm_f32res(a, *blk["out"], x)
            h, _, _, _ = ops.layernorm(x, *blk["ln2"])
            u = ops.gemm(h, *blk["fc"], quick_gelu=True)
            x = ops.gemm_f32res(u, *blk["proj"], x)
        if self.adapter is not None:
            if sequence:
                raise ops._lib.LecbError("the adapter text encoder is only used for prompts (EOT readout), not for sequences")
            xr = x.view(n, l, w)[torch.arange(n, device=x.device), eot_index].contiguous()
            xr2, _, _ = self._adapter_fwd(xr)
            h, _, _, _ = ops.layernorm(xr2, *self.ln_final)
            return ops.gemm(h, self.text_proj_t, out_f32=True)
        h, _, _, _ = ops.layernorm(x, *self.ln_final)
        if sequence:
            return ops.gemm(h, self.text_proj_t, out_f32=True).view(n, l, -1)
        rows = h.view(n, l, w)[torch.arange(n, device=h.device), eot_index].contiguous()
        return ops.gemm(rows, self.text_proj_t, out_f32=True)
