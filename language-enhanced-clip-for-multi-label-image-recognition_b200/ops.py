"""Thin typed wrappers: torch tensors -> raw device pointers + sizes -> C ABI (include/lecb.h).

torch is plumbing here (allocation, streams); every op below runs a hand-written sm_100a kernel."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import EPI_OUT_F32, EPI_QUICKGELU, EPI_RELU, check, lib  # noqa: F401


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need(t, dtype, name):
    if not t.is_cuda:
        raise _lib.LecbError(f"{name} must be a CUDA tensor (lecb200 has no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # every launch goes to the current device's current stream (one process per GPU): a tensor of another device would
        # be dereferenced on the wrong GPU
        raise _lib.LecbError(f"{name} lives on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()} "
                             "(lecb200 runs one process per GPU; wrap the call in torch.cuda.device(...))")
    if t.dtype != dtype:
        raise _lib.LecbError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.LecbError(f"{name} must be contiguous")


def gemm(a, w, bias=None, residual=None, relu=False, quick_gelu=False, out_f32=False, row_sumsq=None, out=None):
    """out[M,N] = epi(a[M,K] @ w[N,K]^T + bias (+ residual)); a, w bf16; bias fp32."""
    _need(a, torch.bfloat16, "a")
    _need(w, torch.bfloat16, "w")
    m, k = a.shape
    n, k2 = w.shape
    assert k == k2, (a.shape, w.shape)
    if bias is not None:
        _need(bias, torch.float32, "bias")
    if residual is not None:
        _need(residual, torch.bfloat16, "residual")
        assert tuple(residual.shape) == (m, n)
    if out is None:
        out = torch.empty((m, n), device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    flags = (EPI_RELU if relu else 0) | (EPI_QUICKGELU if quick_gelu else 0) | (EPI_OUT_F32 if out_f32 else 0)
    check(lib.lecb_gemm_bf16(_ptr(a), _ptr(w), _ptr(bias), _ptr(residual), _ptr(out), _ptr(row_sumsq), m, n, k,
                             flags, _stream()), "lecb_gemm_bf16")
    return out


def gemm_dual(a1, a2, w, bias=None, relu=False, out=None):
    """out[M,N] = epi(a1[M,K1] @ w[:, :K1]^T + a2[M,K2] @ w[:, K1:]^T + bias): the K-concatenated GEMM [a1 | a2] @ w^T without
    the concatenated tensor (lecb_gemm_bf16_dual; the projection-shortcut tail of a Bottleneck, M:46-52).  All bf16, bias fp32,
    K1 and K2 multiples of 64."""
    _need(a1, torch.bfloat16, "a1")
    _need(a2, torch.bfloat16, "a2")
    _need(w, torch.bfloat16, "w")
    m, k1 = a1.shape
    m2, k2 = a2.shape
    n, k = w.shape
    assert m == m2 and k == k1 + k2, (a1.shape, a2.shape, w.shape)
    if bias is not None:
        _need(bias, torch.float32, "bias")
    if out is None:
        out = torch.empty((m, n), device=a1.device, dtype=torch.bfloat16)
    check(lib.lecb_gemm_bf16_dual(_ptr(a1), k1, _ptr(a2), k2, _ptr(w), _ptr(bias), _ptr(out), m, n, EPI_RELU if relu else 0,
                                  _stream()), "lecb_gemm_bf16_dual")
    return out


def gemm_mul_quick_gelu_grad(a, w, v, out=None):
    """out[M,N] = (a[M,K] @ w[N,K]^T) * QuickGELU'(v[M,N]), all bf16: the c_proj data gradient and the QuickGELU backward
    (autograd of M:202-204, 226-227) in one launch; the product is rounded to bf16 once."""
    _need(a, torch.bfloat16, "a")
    _need(w, torch.bfloat16, "w")
    _need(v, torch.bfloat16, "v")
    m, k = a.shape
    n, k2 = w.shape
    assert k == k2 and tuple(v.shape) == (m, n), (a.shape, w.shape, v.shape)
    if out is None:
        out = torch.empty((m, n), device=a.device, dtype=torch.bfloat16)
    check(lib.lecb_gemm_bf16(_ptr(a), _ptr(w), None, _ptr(v), _ptr(out), None, m, n, k, _lib.EPI_MUL_QGELU_GRAD, _stream()),
          "lecb_gemm_bf16")
    return out


def conv3x3(x, w, bias=None, relu=True, out=None, pool=False):
    """x NHWC bf16 [B,H,W,Cin]; w bf16 [Cout,3,3,Cin] (BN folded); -> NHWC bf16 [B,H,W,Cout].
    pool=True appends the 2x2 average pool that follows the conv in the reference (M:147, M:27): fused into the conv
    epilogue when the kernel supports it for this shape (lecb_conv3x3_pool_fusable), else a second kernel."""
    _need(x, torch.bfloat16, "x")
    _need(w, torch.bfloat16, "w")
    b, h, wd, cin = x.shape
    cout = w.shape[0]
    assert tuple(w.shape) == (cout, 3, 3, cin), w.shape
    if bias is not None:
        _need(bias, torch.float32, "bias")
    flags = EPI_RELU if relu else 0
    fused = bool(pool) and lib.lecb_conv3x3_pool_fusable(b, h, wd, cin, cout) == 1
    if fused:
        flags |= _lib.EPI_AVGPOOL2
        out_shape = (b, h // 2, wd // 2, cout)
    else:
        out_shape = (b, h, wd, cout)
    if out is None or pool:
        out = torch.empty(out_shape, device=x.device, dtype=torch.bfloat16)
    check(lib.lecb_conv3x3_bf16(_ptr(x), _ptr(w), _ptr(bias), _ptr(out), b, h, wd, cin, cout, flags, _stream()),
          "lecb_conv3x3_bf16")
    if pool and not fused:
        out = avgpool2x2(out)
    return out


def gemm_f32res(a, w, bias, residual_f32, out=None):
    """fp32 residual stream update: out_f32 = a @ w^T + bias + residual_f32 (text-tower blocks, M:226-227)."""
    _need(a, torch.bfloat16, "a")
    _need(w, torch.bfloat16, "w")
    _need(residual_f32, torch.float32, "residual")
    m, k = a.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((m, n), device=a.device, dtype=torch.float32)
    check(lib.lecb_gemm_bf16(_ptr(a), _ptr(w), _ptr(bias), _ptr(residual_f32), _ptr(out), 0, m, n, k,
                             EPI_OUT_F32 | _lib.EPI_RES_F32, _stream()), "lecb_gemm_bf16")
    return out


def stem_conv1(x, w27, bias):
    """x NCHW fp32 [B,3,H,W]; w27 fp32 [27,Cout]; -> NHWC bf16 [B,H/2,W/2,Cout] (conv+BN+ReLU)."""
    _need(x, torch.float32, "x")
    _need(w27, torch.float32, "w27")
    _need(bias, torch.float32, "bias")
    b, c, h, w = x.shape
    assert c == 3
    cout = w27.shape[1]
    out = torch.empty((b, h // 2, w // 2, cout), device=x.device, dtype=torch.bfloat16)
    check(lib.lecb_stem_conv1(_ptr(x), _ptr(w27), _ptr(bias), _ptr(out), b, h, w, cout, _stream()), "lecb_stem_conv1")
    return out


def avgpool2x2(x):
    _need(x, torch.bfloat16, "x")
    b, h, w, c = x.shape
    out = torch.empty((b, h // 2, w // 2, c), device=x.device, dtype=torch.bfloat16)
    check(lib.lecb_avgpool2x2(_ptr(x), _ptr(out), b, h, w, c, _stream()), "lecb_avgpool2x2")
    return out


def token_mean(x, want_f32=False):
    """x bf16 [B,P,C] -> bf16 [B,C] (and fp32 copy if asked)."""
    _need(x, torch.bfloat16, "x")
    b, p, c = x.shape
    ob = torch.empty((b, c), device=x.device, dtype=torch.bfloat16)
    of = torch.empty((b, c), device=x.device, dtype=torch.float32) if want_f32 else None
    check(lib.lecb_token_mean(_ptr(x), _ptr(ob), _ptr(of), b, p, c, _stream()), "lecb_token_mean")
    return (ob, of) if want_f32 else ob


def l2norm_rows(x, out_dtype=None):
    """y = x / ||x||_2 along the last dim (no epsilon, like the reference)."""
    assert x.dtype in (torch.float32, torch.bfloat16) and x.is_cuda and x.is_contiguous()
    out_dtype = out_dtype or x.dtype
    d = x.shape[-1]
    rows = x.numel() // d
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    check(lib.lecb_l2norm_rows(_ptr(x), _ptr(y), rows, d, int(x.dtype == torch.bfloat16),
                               int(out_dtype == torch.bfloat16), _stream()), "lecb_l2norm_rows")
    return y


def layernorm(x, gamma, beta, eps=1e-5, out_bf16=True, out_f32=False, save_stats=False):
    _need(x, torch.float32, "x")
    d = x.shape[-1]
    rows = x.numel() // d
    yb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if out_bf16 else None
    yf = torch.empty(x.shape, device=x.device, dtype=torch.float32) if out_f32 else None
    mean = torch.empty((rows,), device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty((rows,), device=x.device, dtype=torch.float32) if save_stats else None
    check(lib.lecb_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(yb), _ptr(yf), _ptr(mean), _ptr(rstd), rows, d,
                                 float(eps), _stream()), "lecb_layernorm_fwd")
    return yb, yf, mean, rstd


def attnpool_query0(q, kmat, vmat, b, p, heads):
    _need(q, torch.float32, "q")
    _need(kmat, torch.bfloat16, "kmat")
    _need(vmat, torch.bfloat16, "vmat")
    c = q.shape[-1]
    out = torch.empty((b, c), device=q.device, dtype=torch.bfloat16)
    check(lib.lecb_attnpool_query0(_ptr(q), _ptr(kmat), _ptr(vmat), _ptr(out), b, p, c, heads, _stream()),
          "lecb_attnpool_query0")
    return out


def causal_attn(qkv, n, l, w, heads):
    """Text-tower masked self-attention (M:221-223, mask M:364-370) on the tcgen05 attention kernel."""
    return attn_fwd(qkv, n, l, w, heads, causal=True)


def causal_attn_smem(qkv, n, l, w, heads):
    """The CUDA-core shared-memory variant (kept as an independent cross-check of the tensor-core kernel)."""
    _need(qkv, torch.bfloat16, "qkv")
    out = torch.empty((n * l, w), device=qkv.device, dtype=torch.bfloat16)
    check(lib.lecb_causal_attn_fwd(_ptr(qkv), _ptr(out), n, l, w, heads, _stream()), "lecb_causal_attn_fwd")
    return out


def attn_fwd(qkv, b, t, w, heads, q_rows=None, causal=False, out=None):
    """Multi-head attention on tcgen05: qkv bf16 [B*T,3W] -> bf16 [B*T,W] (rows >= q_rows of each sequence untouched)."""
    _need(qkv, torch.bfloat16, "qkv")
    assert tuple(qkv.shape) == (b * t, 3 * w), (qkv.shape, b, t, w)
    if out is None:
        out = torch.empty((b * t, w), device=qkv.device, dtype=torch.bfloat16)
    check(lib.lecb_attn_fwd(_ptr(qkv), _ptr(out), b, t, w, heads, t if q_rows is None else q_rows, int(causal),
                            _stream()), "lecb_attn_fwd")
    return out


def patchify(x, patch, kpad):
    """x NCHW fp32 [B,3,H,W] -> bf16 [B*(H/p)*(W/p), kpad] patch rows (A operand of the patch-embedding GEMM)."""
    _need(x, torch.float32, "x")
    b, c, h, w = x.shape
    assert c == 3
    out = torch.empty((b * (h // patch) * (w // patch), kpad), device=x.device, dtype=torch.bfloat16)
    check(lib.lecb_patchify(_ptr(x), _ptr(out), b, h, w, patch, kpad, _stream()), "lecb_patchify")
    return out


def vit_embed_ln(emb, cls, pos, gamma, beta, b, t, eps=1e-5):
    """fp32 [B*T, D] = ln_pre([class_embedding ; emb] + positional_embedding)  (M:262-266)."""
    _need(emb, torch.bfloat16, "emb")
    d = emb.shape[-1]
    out = torch.empty((b * t, d), device=emb.device, dtype=torch.float32)
    check(lib.lecb_vit_embed_ln(_ptr(emb), _ptr(cls), _ptr(pos), _ptr(gamma), _ptr(beta), _ptr(out), b, t, d,
                                float(eps), _stream()), "lecb_vit_embed_ln")
    return out


def copy_cols(src, col0, cols, out=None):
    """bf16 [rows, cols] = src[:, col0:col0+cols] (hand-written strided copy)."""
    _need(src, torch.bfloat16, "src")
    rows, ld = src.shape
    if out is None:
        out = torch.empty((rows, cols), device=src.device, dtype=torch.bfloat16)
    check(lib.lecb_copy_cols(_ptr(src), ld, col0, _ptr(out), out.shape[1], rows, cols, _stream()), "lecb_copy_cols")
    return out


def head_aggregate(dots, b, p, k, n_txt, row_sumsq=None, row_mask=None, logit_scale=4.0, spatial_scale=50.0,
                   want_maps=True):
    """dots fp32 [B*P, ldn] -> logits_local [B,K] (+ neg_map, pos_map [P,B,K])."""
    _need(dots, torch.float32, "dots")
    ldn = dots.shape[-1]
    out = torch.empty((b, k), device=dots.device, dtype=torch.float32)
    # rows of masked tokens are not written by the kernel: with a mask the maps start from zeros
    alloc = torch.zeros if row_mask is not None else torch.empty
    neg = alloc((p, b, k), device=dots.device, dtype=torch.float32) if want_maps else None
    pos = alloc((p, b, k), device=dots.device, dtype=torch.float32) if want_maps else None
    if row_mask is not None:
        _need(row_mask, torch.uint8, "row_mask")
    check(lib.lecb_head_aggregate(_ptr(dots), ldn, _ptr(row_sumsq), _ptr(row_mask), _ptr(out), _ptr(neg), _ptr(pos),
                                  b, p, k, n_txt, float(logit_scale), float(spatial_scale), _stream()),
          "lecb_head_aggregate")
    return out, neg, pos


def global_logits(g_unit, tpos, g_add=None, scale=4.0):
    _need(g_unit, torch.float32, "g_unit")
    _need(tpos, torch.float32, "tpos")
    b, d = g_unit.shape
    k = tpos.shape[0]
    out = torch.empty((b, k), device=g_unit.device, dtype=torch.float32)
    check(lib.lecb_global_logits(_ptr(g_unit), _ptr(g_add), _ptr(tpos), _ptr(out), b, d, k, float(scale), _stream()),
          "lecb_global_logits")
    return out


def asl_fwd_bwd(logits, targets, gamma_neg=2.0, gamma_pos=1.0, clip=0.05, eps=1e-8, thresh_pos=0.9, thresh_neg=0.9,
                partial=False, want_grad=True):
    _need(logits, torch.float32, "logits")
    _need(targets, torch.float32, "targets")
    b, k = logits.shape
    grad = torch.empty_like(logits) if want_grad else None
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    check(lib.lecb_asl_fwd_bwd(_ptr(logits), _ptr(targets), _ptr(grad), _ptr(loss), b, k, gamma_neg, gamma_pos, clip,
                               eps, thresh_pos, thresh_neg, int(partial), _stream()), "lecb_asl_fwd_bwd")
    return loss, grad


def ranking_fwd_bwd(logits, targets, scale=2.0, margin=1.0, want_grad=True):
    _need(logits, torch.float32, "logits")
    _need(targets, torch.float32, "targets")
    b, k = logits.shape
    grad = torch.empty_like(logits) if want_grad else None
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    check(lib.lecb_ranking_fwd_bwd(_ptr(logits), _ptr(targets), _ptr(grad), _ptr(loss), b, k, float(scale),
                                   float(margin), _stream()), "lecb_ranking_fwd_bwd")
    return loss, grad


# ---------------------------------------------------------------------------------------------
# prompt-tuning backward ops
# ---------------------------------------------------------------------------------------------
def quick_gelu_fwd(v):
    _need(v, torch.bfloat16, "v")
    u = torch.empty_like(v)
    check(lib.lecb_quick_gelu_fwd(_ptr(v), _ptr(u), v.numel(), _stream()), "lecb_quick_gelu_fwd")
    return u


def quick_gelu_bwd(du, v):
    _need(du, torch.bfloat16, "du")
    _need(v, torch.bfloat16, "v")
    dv = torch.empty_like(v)
    check(lib.lecb_quick_gelu_bwd(_ptr(du), _ptr(v), _ptr(dv), v.numel(), _stream()), "lecb_quick_gelu_bwd")
    return dv


def layernorm_bwd(dy, x, gamma, mean, rstd, dx_in=None, want_bf16=True):
    """-> (dx fp32, dx bf16|None) with dx = dx_in + LN'(dy)."""
    _need(dy, torch.float32, "dy")
    _need(x, torch.float32, "x")
    d = x.shape[-1]
    rows = x.numel() // d
    dxf = torch.empty_like(x)
    dxb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    check(lib.lecb_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd), _ptr(dx_in), _ptr(dxf),
                                 _ptr(dxb), rows, d, _stream()), "lecb_layernorm_bwd")
    return dxf, dxb


def causal_attn_bwd(qkv, dout, n, l, w, heads):
    """Backward of the text tower's masked attention on tcgen05 (dqkv bf16 [N*L,3W])."""
    _need(qkv, torch.bfloat16, "qkv")
    _need(dout, torch.bfloat16, "dout")
    dqkv = torch.empty_like(qkv)
    check(lib.lecb_attn_causal_bwd(_ptr(qkv), _ptr(dout), _ptr(dqkv), n, l, w, heads, _stream()), "lecb_attn_causal_bwd")
    return dqkv


def causal_attn_bwd_smem(qkv, dout, n, l, w, heads):
    """CUDA-core shared-memory variant (fp32 probabilities), kept as an independent cross-check."""
    _need(qkv, torch.bfloat16, "qkv")
    _need(dout, torch.bfloat16, "dout")
    dqkv = torch.empty_like(qkv)
    check(lib.lecb_causal_attn_bwd(_ptr(qkv), _ptr(dout), _ptr(dqkv), n, l, w, heads, _stream()), "lecb_causal_attn_bwd")
    return dqkv


def l2norm_bwd(x, dy):
    _need(x, torch.float32, "x")
    _need(dy, torch.float32, "dy")
    d = x.shape[-1]
    dx = torch.empty_like(x)
    check(lib.lecb_l2norm_bwd(_ptr(x), _ptr(dy), _ptr(dx), x.numel() // d, d, _stream()), "lecb_l2norm_bwd")
    return dx


def head_aggregate_bwd(dots, grad_local, b, p, k, n_txt, row_sumsq=None, row_mask=None, logit_scale=4.0,
                       spatial_scale=50.0):
    _need(dots, torch.float32, "dots")
    _need(grad_local, torch.float32, "grad_local")
    d_dots = torch.empty_like(dots)
    if dots.shape[-1] > n_txt * k:
        d_dots.zero_()
    check(lib.lecb_head_aggregate_bwd(_ptr(dots), dots.shape[-1], _ptr(row_sumsq), _ptr(row_mask), _ptr(grad_local),
                                      _ptr(d_dots), b, p, k, n_txt, float(logit_scale), float(spatial_scale), _stream()),
          "lecb_head_aggregate_bwd")
    return d_dots


def tn_gemm_small(a, b, j, out=None, alpha=1.0, accumulate=False):
    """out[J,D] (+)= alpha * a[:, :J]^T @ b;  a fp32 [R,lda], b bf16|fp32 [R,D]."""
    _need(a, torch.float32, "a")
    assert b.is_cuda and b.is_contiguous() and b.dtype in (torch.bfloat16, torch.float32)
    r, lda = a.shape
    d = b.shape[1]
    assert b.shape[0] == r
    if out is None:
        out = torch.empty((j, d), device=a.device, dtype=torch.float32)
        assert not accumulate
    check(lib.lecb_tn_gemm_small(_ptr(a), lda, _ptr(b), int(b.dtype == torch.bfloat16), _ptr(out), r, j, d, float(alpha),
                                 int(accumulate), _stream()), "lecb_tn_gemm_small")
    return out


# ---------------------------------------------------------------------------------------------
# round 2: uint8 stem input, remaining losses, multi-tensor updates
# ---------------------------------------------------------------------------------------------
CLIP_PIXEL_MEAN = (0.48145466, 0.4578275, 0.40821073)      # cfg.INPUT.PIXEL_MEAN / PIXEL_STD of every shipped yaml
CLIP_PIXEL_STD = (0.26862954, 0.26130258, 0.27577711)


def _f3(vals):
    import ctypes
    return (ctypes.c_float * 3)(*[float(v) for v in vals])


def stem_conv1_u8(x, w27, bias, mean=CLIP_PIXEL_MEAN, std=CLIP_PIXEL_STD):
    """x NHWC uint8 [B,H,W,3] raw pixels; ToTensor + Normalize(mean, std) applied inside the kernel;
    -> NHWC bf16 [B,H/2,W/2,Cout] (conv + folded BN + ReLU), identical to stem_conv1 on the normalised float tensor."""
    _need(x, torch.uint8, "x")
    _need(w27, torch.float32, "w27")
    _need(bias, torch.float32, "bias")
    b, h, w, c = x.shape
    assert c == 3
    cout = w27.shape[1]
    out = torch.empty((b, h // 2, w // 2, cout), device=x.device, dtype=torch.bfloat16)
    m, s = _f3(mean), _f3(std)
    check(lib.lecb_stem_conv1_u8(_ptr(x), _ptr(w27), _ptr(bias), m, s, _ptr(out), b, h, w, cout, _stream()),
          "lecb_stem_conv1_u8")
    return out


def ranking_cooc_fwd_bwd(logits, targets, pair_weights, scale=2.0, margin=1.0, want_grad=True):
    _need(logits, torch.float32, "logits")
    _need(targets, torch.float32, "targets")
    _need(pair_weights, torch.float32, "pair_weights")
    b, k = logits.shape
    assert tuple(pair_weights.shape) == (k, k)
    grad = torch.empty_like(logits) if want_grad else None
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    check(lib.lecb_ranking_cooc_fwd_bwd(_ptr(logits), _ptr(targets), _ptr(pair_weights), _ptr(grad), _ptr(loss), b, k,
                                        float(scale), float(margin), _stream()), "lecb_ranking_cooc_fwd_bwd")
    return loss, grad


def kl_softmax_fwd_bwd(logits, logits_target, weight=1.0, want_grad=True):
    """weight * KLDivLoss(batchmean)(log_softmax(logits), softmax(logits_target)) and its gradient w.r.t. logits."""
    _need(logits, torch.float32, "logits")
    _need(logits_target, torch.float32, "logits_target")
    assert logits.shape == logits_target.shape
    b, k = logits.shape
    grad = torch.empty_like(logits) if want_grad else None
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    check(lib.lecb_kl_softmax_fwd_bwd(_ptr(logits), _ptr(logits_target), _ptr(grad), _ptr(loss), b, k, float(weight),
                                      _stream()), "lecb_kl_softmax_fwd_bwd")
    return loss, grad


class TensorList:
    """HOST arrays of device pointers + element counts for the multi-tensor kernels (built once per parameter list)."""

    def __init__(self, tensors, allow_none=False):
        import ctypes
        self.n = len(tensors)
        self.tensors = list(tensors)
        for t in tensors:
            if t is None:
                assert allow_none
                continue
            _need(t, torch.float32, "tensor")
        self.ptrs = (ctypes.c_void_p * self.n)(*[None if t is None else t.data_ptr() for t in tensors])
        self.sizes = None

    def with_sizes(self, sizes):
        import ctypes
        self.sizes = (ctypes.c_longlong * self.n)(*[int(s) for s in sizes])
        return self


def _sizes(tensors):
    import ctypes
    return (ctypes.c_longlong * len(tensors))(*[int(t.numel()) for t in tensors])


def ema_update(live, twin, momentum):
    """twin_i <- momentum * twin_i + (1 - momentum) * live_i for every pair, one launch (T:554-559)."""
    a, b = TensorList(live), TensorList(twin)
    check(lib.lecb_ema_update(a.ptrs, b.ptrs, _sizes(live), a.n, float(momentum), float(1.0 - float(momentum)), _stream()),
          "lecb_ema_update")


def pack_f32(tensors, like, flat=None):
    """flat fp32 <- concat of `tensors` (None entries contribute zeros of the size of the matching `like` tensor)."""
    total = sum(int(t.numel()) for t in like)
    if flat is None:
        flat = torch.empty((total,), device=like[0].device, dtype=torch.float32)
    src = TensorList(tensors, allow_none=True)
    check(lib.lecb_pack_f32(src.ptrs, _sizes(like), src.n, _ptr(flat), _stream()), "lecb_pack_f32")
    return flat


def unpack_scale_f32(flat, tensors, scale=1.0):
    _need(flat, torch.float32, "flat")
    dst = TensorList(tensors)
    check(lib.lecb_unpack_scale_f32(_ptr(flat), dst.ptrs, _sizes(tensors), dst.n, float(scale), _stream()),
          "lecb_unpack_scale_f32")


def sgd_step(flat_grad, params, momentum_bufs, lr, momentum=0.0, weight_decay=0.0, grad_scale=1.0):
    """torch.optim.SGD step (dampening 0, no nesterov, zero-initialised momentum buffers) from a flat gradient."""
    _need(flat_grad, torch.float32, "flat_grad")
    p = TensorList(params)
    m = TensorList(momentum_bufs) if momentum_bufs is not None else None
    check(lib.lecb_sgd_step(_ptr(flat_grad), p.ptrs, m.ptrs if m is not None else None, _sizes(params), p.n,
                            float(grad_scale), float(lr), float(momentum), float(weight_decay), _stream()), "lecb_sgd_step")


def resample_bce_fwd_bwd(logits, labels, freq_inv=None, init_bias=None, map_alpha=10.0, map_beta=0.2, map_gamma=0.1,
                         neg_scale=0.0, focal=False, focal_gamma=2.0, balance_param=2.0, loss_weight=1.0, want_grad=True):
    """ResampleLoss (dbl.py:263-445, use_sigmoid) forward + backward; freq_inv / init_bias fp32 [K] or None."""
    _need(logits, torch.float32, "logits")
    _need(labels, torch.float32, "labels")
    b, k = logits.shape
    for t, nm in ((freq_inv, "freq_inv"), (init_bias, "init_bias")):
        if t is not None:
            _need(t, torch.float32, nm)
            assert t.numel() == k
    grad = torch.empty_like(logits) if want_grad else None
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    scratch = torch.empty((2,), device=logits.device, dtype=torch.float32) if focal else None
    check(lib.lecb_resample_bce_fwd_bwd(_ptr(logits), _ptr(labels), _ptr(freq_inv), _ptr(init_bias), _ptr(grad), _ptr(loss),
                                        _ptr(scratch), b, k, float(map_alpha), float(map_beta), float(map_gamma), float(neg_scale),
                                        int(bool(focal)), float(focal_gamma), float(balance_param), float(loss_weight), _stream()),
          "lecb_resample_bce_fwd_bwd")
    return loss, grad


def residual_relu_fwd(x, z):
    """out = x + relu(z) on fp32 (the residual adapter, Caption_distill_double_adapter.py:109)."""
    _need(x, torch.float32, "x")
    _need(z, torch.float32, "z")
    out = torch.empty_like(x)
    check(lib.lecb_residual_relu_fwd(_ptr(x), _ptr(z), _ptr(out), x.numel(), _stream()), "lecb_residual_relu_fwd")
    return out


def relu_bwd(dy, z):
    """bf16 dz = dy * 1[z > 0]; dy fp32, z fp32 (pre-activation) or bf16 (post-activation)."""
    _need(dy, torch.float32, "dy")
    assert z.is_cuda and z.is_contiguous() and z.dtype in (torch.float32, torch.bfloat16) and z.numel() == dy.numel()
    dz = torch.empty(dy.shape, device=dy.device, dtype=torch.bfloat16)
    check(lib.lecb_relu_bwd(_ptr(dy), _ptr(z), int(z.dtype == torch.bfloat16), _ptr(dz), dy.numel(), _stream()), "lecb_relu_bwd")
    return dz
