"""Caption retrieval (reference T:444-448) on the B200.

Fused path (default): the fp32 query is split into an exact fp16 (hi, lo) pair laid out as one [B, 2D] operand, ONE
tcgen05 GEMM multiplies both halves against the fp16 bank (each bank tile staged once per query block) and keeps the ten
largest similarities of every row in registers; a merge kernel reduces the per-CTA lists and a gather kernel averages
the selected bank rows.  The [B, N] fp32 similarity matrix of T:445 is never written (225 MB at 256 x 220 000).
`retrieve_mean_unfused` is round 1's explicit-matrix path, kept as the cross-check."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check, lib
from .ops import _ptr, _stream

_MAX_SLOTS = 160          # >= the GEMM grid (one CTA per SM)


def _check_bank(g_unit, bank):
    if bank.dtype != torch.float16:
        raise _lib.LecbError("caption bank must be fp16 (what the reference's feature builder stores)")
    if not (bank.is_cuda and bank.is_contiguous()):
        raise _lib.LecbError("caption bank must be a contiguous CUDA tensor")
    b, d = g_unit.shape
    if bank.shape[1] != d or d % 64 != 0 or bank.shape[0] < 10:
        raise _lib.LecbError(f"bank shape {tuple(bank.shape)} incompatible with features of dim {d} (D % 64, N >= 10)")


def retrieve_mean(g_unit: torch.Tensor, bank: torch.Tensor):
    """g_unit fp32 [B,D] unit rows; bank fp16 [N,D], any N >= 10 (the file generate_caption_text_features.py writes).
    -> (mean of the top-10 bank rows, fp32 [B,D], rounded through fp16 like the reference; top-10 scores [B,10])."""
    _check_bank(g_unit, bank)
    b, d = g_unit.shape
    n = bank.shape[0]
    dev = g_unit.device
    hilo = torch.empty((b, 2 * d), device=dev, dtype=torch.float16)
    check(lib.lecb_split_f16_hilo(_ptr(g_unit), _ptr(hilo), b, d, _stream()), "lecb_split_f16_hilo")
    part_val = torch.empty((b, _MAX_SLOTS, 10), device=dev, dtype=torch.float32)
    part_idx = torch.empty((b, _MAX_SLOTS, 10), device=dev, dtype=torch.int32)
    used = ctypes.c_int(0)
    check(lib.lecb_gemm_topk10(_ptr(hilo), _ptr(bank), b, n, d, _ptr(part_val), _ptr(part_idx), _MAX_SLOTS,
                               ctypes.byref(used), _stream()), "lecb_gemm_topk10")
    vals = torch.empty((b, 10), device=dev, dtype=torch.float32)
    idx = torch.empty((b, 10), device=dev, dtype=torch.int32)
    check(lib.lecb_topk10_merge(_ptr(part_val), _ptr(part_idx), _MAX_SLOTS, b, _ptr(vals), _ptr(idx), _stream()),
          "lecb_topk10_merge")
    g_add = torch.empty((b, d), device=dev, dtype=torch.float32)
    check(lib.lecb_gather_mean10(_ptr(bank), 1, _ptr(idx), _ptr(g_add), b, d, _stream()), "lecb_gather_mean10")
    return g_add, vals


def retrieve_mean_unfused(g_unit: torch.Tensor, bank: torch.Tensor, return_idx=False):
    """Round-1 path: two fp16 GEMMs into an explicit [B, N] similarity matrix + lecb_topk10 (N % 8 == 0)."""
    _check_bank(g_unit, bank)
    b, d = g_unit.shape
    n = bank.shape[0]
    if n % 8 != 0:
        raise _lib.LecbError("retrieve_mean_unfused needs N % 8 == 0 (use retrieve_mean)")
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
    return (g_add, vals, idx) if return_idx else (g_add, vals)
