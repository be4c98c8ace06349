"""Caption retrieval (reference T:444-448) on the B200: similarity GEMM on tcgen05 with an exact
fp16 (hi, lo) split of the fp32 query, per-row top-10, gather + mean of the selected bank rows."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, lib
from .ops import _ptr, _stream


def retrieve_mean(g_unit: torch.Tensor, bank: torch.Tensor):
    """g_unit fp32 [B,D] unit rows; bank fp16 [N,D] (the format generate_caption_text_features.py writes).
    -> (mean of the top-10 bank rows, fp32 [B,D], rounded through fp16 like the reference; top-10 scores [B,10])."""
    if bank.dtype != torch.float16:
        raise _lib.LecbError("caption bank must be fp16 (what the reference's feature builder stores)")
    if not (bank.is_cuda and bank.is_contiguous()):
        raise _lib.LecbError("caption bank must be a contiguous CUDA tensor")
    b, d = g_unit.shape
    n = bank.shape[0]
    if bank.shape[1] != d or n % 8 != 0 or d % 32 != 0:
        raise _lib.LecbError(f"bank shape {tuple(bank.shape)} incompatible with features of dim {d} (N % 8, D % 32)")
    dev = g_unit.device
    hi = torch.empty((b, d), device=dev, dtype=torch.float16)
    lo = torch.empty((b, d), device=dev, dtype=torch.float16)
    check(lib.lecb_split_f16(_ptr(g_unit), _ptr(hi), _ptr(lo), b * d, _stream()), "lecb_split_f16")
    sim = torch.empty((b, n), device=dev, dtype=torch.float32)
    f = _lib.EPI_OUT_F32 | _lib.GEMM_F16_OPERANDS
    check(lib.lecb_gemm_bf16(_ptr(hi), _ptr(bank), 0, 0, _ptr(sim), 0, b, n, d, f, _stream()), "lecb_gemm_bf16")
    check(lib.lecb_gemm_bf16(_ptr(lo), _ptr(bank), 0, _ptr(sim), _ptr(sim), 0, b, n, d, f | _lib.EPI_RES_F32, _stream()),
          "lecb_gemm_bf16")
    vals = torch.empty((b, 10), device=dev, dtype=torch.float32)
    idx = torch.empty((b, 10), device=dev, dtype=torch.int32)
    check(lib.lecb_topk10(_ptr(sim), n, b, n, _ptr(vals), _ptr(idx), _stream()), "lecb_topk10")
    g_add = torch.empty((b, d), device=dev, dtype=torch.float32)
    check(lib.lecb_gather_mean10(_ptr(bank), 1, _ptr(idx), _ptr(g_add), b, d, _stream()), "lecb_gather_mean10")
    return g_add, vals
