"""Loss functions with the reference signatures (trainers/utils.py, U) backed by fused fwd+bwd kernels.

`ranking_loss(y_pred, y_true, scale_=2.0, margin_=1)`  U:85-93  (does NOT scale y_pred in place, unlike U:86)
`ASL_loss(inputs, targets)`                             U:184-190
`dualcoop_loss(inputs, inputs_g, targets)`              U:175-181
`AsymmetricLoss_partial(...)`                           U:126-173
`ranking_loss_with_cooccurrence(y_pred, y_true, cooccurrence, scale_=2.0, margin_=1)`   U:95-110 (T:842-850)
`ema_consistency_loss(output, output_m, output_local, output_local_m)`                   the two KL terms of T:809-813
Each launch produces the scalar loss and dloss/dlogits together; autograd just scales the stored gradient."""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, fn):
        loss, grad = fn(logits.detach().float().contiguous())
        ctx.save_for_backward(grad)
        ctx.in_dtype = logits.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).to(ctx.in_dtype), None


def ranking_loss(y_pred, y_true, scale_=2.0, margin_=1):
    t = y_true.detach().float().contiguous()
    return _FusedLoss.apply(y_pred, lambda x: ops.ranking_fwd_bwd(x, t, scale_, margin_))


def cooccurrence_pair_weights(cooccurrence):
    """U:99-102: log(1 / (p + 1e-6)), zero diagonal, every row divided by its mean -> fp32 [K,K] (tiny; plain torch)."""
    w = (1 / (cooccurrence.float() + 1e-6)).log()
    w = w * (1 - torch.eye(w.shape[0], w.shape[1], device=w.device))
    return (w / w.mean(-1)[:, None]).contiguous()


def ranking_loss_with_cooccurrence(y_pred, y_true, cooccurrence, scale_=2.0, margin_=1):
    t = y_true.detach().float().contiguous()
    w = cooccurrence_pair_weights(cooccurrence.detach())
    return _FusedLoss.apply(y_pred, lambda x: ops.ranking_cooc_fwd_bwd(x, t, w, scale_, margin_))


def kl_softmax(output, output_m, weight=1.0):
    """weight * nn.KLDivLoss(reduction="batchmean")(F.log_softmax(output, -1), F.softmax(output_m, -1)); gradient to
    `output` only (the reference computes `output_m` under torch.no_grad(), T:517)."""
    tgt = output_m.detach().float().contiguous()
    return _FusedLoss.apply(output, lambda x: ops.kl_softmax_fwd_bwd(x, tgt, weight))


def ema_consistency_loss(output, output_m, output_local, output_local_m, local_weight=10000.0):
    """`ema_loss` of T:809-811: KL on the global logits + 10000 x KL on the local logits."""
    return kl_softmax(output, output_m) + kl_softmax(output_local, output_local_m, local_weight)


class AsymmetricLoss_partial(nn.Module):
    def __init__(self, gamma_neg=4, gamma_pos=1, clip=0.05, eps=1e-8, disable_torch_grad_focal_loss=True):
        super().__init__()
        if not disable_torch_grad_focal_loss:
            raise NotImplementedError("lecb200 implements the reference's default: focal weight excluded from autograd (U:162-170)")
        self.gamma_neg, self.gamma_pos, self.clip, self.eps = gamma_neg, gamma_pos, clip, eps

    def forward(self, x, y, thresh_pos=0.9, thresh_neg=-0.9, if_partial=True):
        t = y.detach().float().contiguous()
        clip = 0.0 if self.clip is None else float(self.clip)
        return _FusedLoss.apply(x, lambda z: ops.asl_fwd_bwd(z, t, float(self.gamma_neg), float(self.gamma_pos), clip,
                                                             float(self.eps), float(thresh_pos), float(thresh_neg),
                                                             bool(if_partial)))


def dualcoop_loss(inputs, inputs_g, targets):
    return AsymmetricLoss_partial(gamma_neg=2, gamma_pos=1, clip=0.05)(inputs, targets, thresh_pos=0.9, thresh_neg=-0.9)


def ASL_loss(inputs, targets):
    return AsymmetricLoss_partial(gamma_neg=2, gamma_pos=1, clip=0.05)(inputs, targets, thresh_pos=0.9, thresh_neg=0.9,
                                                                         if_partial=False)


class ResampleLoss(nn.Module):
    """`ResampleLoss` of trainers/dbl.py:263-445 (LOSSFUNC 'dbl', built at T:818-830) on one fused kernel.

    Same constructor keywords as the reference; the class frequencies come from `freq_file` (a pickle with `class_freq` /
    `neg_class_freq`, what `mmcv.load` reads there) or directly from the additive `class_freq=` / `neg_class_freq=` arguments.
    Supported: use_sigmoid=True, partial=False, reweight_func None | 'rebalance', weight_norm=None, any focal / map_param /
    logit_reg setting — the combinations T:823-839 construct.  The in-place `logits += init_bias` side effect of dbl.py:405
    is not reproduced (the argument is left untouched, like `ranking_loss`)."""

    def __init__(self, use_sigmoid=False, reduction='mean', loss_weight=1.0, partial=False,
                 focal=dict(focal=True, balance_param=2.0, gamma=2), CB_loss=dict(CB_beta=0.9, CB_mode='average_w'),
                 map_param=dict(alpha=10.0, beta=0.2, gamma=0.1), logit_reg=dict(neg_scale=5.0, init_bias=0.1),
                 reweight_func=None, weight_norm=None, freq_file='./class_freq.pkl', class_freq=None, neg_class_freq=None,
                 device="cuda"):
        super().__init__()
        if not use_sigmoid or partial:
            raise NotImplementedError("lecb200 ResampleLoss: use_sigmoid=True, partial=False (the configuration built at T:823-830)")
        if reweight_func not in (None, "rebalance") or weight_norm is not None:
            raise NotImplementedError(f"lecb200 ResampleLoss: reweight_func={reweight_func!r} / weight_norm={weight_norm!r} is not built")
        if reduction != "mean":
            raise NotImplementedError("lecb200 ResampleLoss: reduction='mean' (dbl.py:61-63 averages whatever is asked)")
        if class_freq is None:
            import pickle
            with open(freq_file, "rb") as f:
                stats = pickle.load(f)
            class_freq, neg_class_freq = stats["class_freq"], stats["neg_class_freq"]
        cf = torch.as_tensor(class_freq, dtype=torch.float32, device=device)
        ncf = torch.as_tensor(neg_class_freq, dtype=torch.float32, device=device)
        self.loss_weight, self.reweight_func = float(loss_weight), reweight_func
        self.focal, self.gamma, self.balance_param = bool(focal["focal"]), float(focal["gamma"]), float(focal["balance_param"])
        self.map_alpha, self.map_beta, self.map_gamma = float(map_param["alpha"]), float(map_param["beta"]), float(map_param["gamma"])
        self.logit_reg = dict(logit_reg)
        self.neg_scale = float(logit_reg["neg_scale"]) if "neg_scale" in logit_reg else 1.0
        init_bias = float(logit_reg["init_bias"]) if "init_bias" in logit_reg else 0.0
        train_num = cf[0] + ncf[0]
        self.register_buffer("class_freq", cf)
        self.register_buffer("freq_inv", (torch.ones_like(cf) / cf).contiguous())                        # dbl.py:341
        self.register_buffer("init_bias", (-torch.log(train_num / cf - 1) * init_bias / self.neg_scale).contiguous())      # dbl.py:338-339

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        y = label.detach().float().contiguous()
        fi = self.freq_inv if self.reweight_func == "rebalance" else None
        ib = self.init_bias if "init_bias" in self.logit_reg else None
        ns = self.neg_scale if "neg_scale" in self.logit_reg else 0.0
        return _FusedLoss.apply(cls_score, lambda x: ops.resample_bce_fwd_bwd(
            x, y, fi, ib, self.map_alpha, self.map_beta, self.map_gamma, ns, self.focal, self.gamma, self.balance_param,
            self.loss_weight))
