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


def conv3x3(x, w, bias=None, relu=True, out=None):
    """x NHWC bf16 [B,H,W,Cin]; w bf16 [Cout,3,3,Cin] (BN folded); -> NHWC bf16 [B,H,W,Cout]."""
    _need(x, torch.bfloat16, "x")
    _need(w, torch.bfloat16, "w")
    b, h, wd, cin = x.shape
    cout = w.shape[0]
    assert tuple(w.shape) == (cout, 3, 3, cin), w.shape
    if bias is not None:
        _need(bias, torch.float32, "bias")
    if out is None:
        out = torch.empty((b, h, wd, cout), device=x.device, dtype=torch.bfloat16)
    check(lib.lecb_conv3x3_bf16(_ptr(x), _ptr(w), _ptr(bias), _ptr(out), b, h, wd, cin, cout,
                                EPI_RELU if relu else 0, _stream()), "lecb_conv3x3_bf16")
    return out
