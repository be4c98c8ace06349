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
