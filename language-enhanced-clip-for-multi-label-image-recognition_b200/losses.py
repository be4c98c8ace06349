"""Loss functions with the reference signatures (trainers/utils.py, U) backed by fused fwd+bwd kernels.

`ranking_loss(y_pred, y_true, scale_=2.0, margin_=1)`  U:85-93  (does NOT scale y_pred in place, unlike U:86)
`ASL_loss(inputs, targets)`                             U:184-190
`dualcoop_loss(inputs, inputs_g, targets)`              U:175-181
`AsymmetricLoss_partial(...)`                           U:126-173
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
