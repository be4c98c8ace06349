"""Test-time score fusion around the scoring path, on the GPU.

Mirrors, with the reference's names and argument meaning:
  * `adjust_predictions(raw_predictions, normalized_cooccurrence_matrix, weight)` — the nested helper of
    `Caption_distill_double.test` (T:611-615) and `normalized_cooccurrence(adj, nums)` (T:632-634);
  * `aggregate_blocks(output, output_blocks, threshold=0.3, weight=1.4)` — the inline max / min / threshold
    aggregation of the sliding-window scores (T:655-662);
  * `fuse(data, sims_scores, threshold=0.2)` / `fuse6(...)` — gen_final_ans.py:18-71 (the reference reads
    `sims_scores` from a module global; here it is an argument).
Every function is one hand-written kernel launch (csrc/postprocess.cu); no CPU path."""
from __future__ import annotations

import torch

from ._lib import check, lib
from .ops import _need, _ptr, _stream


def normalized_cooccurrence(adj, nums, device="cuda"):
    """T:632-634: p = adj / nums[:, None]; p = p / p.sum(-1)[:, None]  (tiny [K,K] setup, done once)."""
    p = torch.as_tensor(adj, dtype=torch.float32, device=device) / torch.as_tensor(nums, dtype=torch.float32, device=device)[:, None]
    return (p / p.sum(-1)[:, None]).contiguous()


def adjust_predictions(raw_predictions, normalized_cooccurrence_matrix, weight=1.0):
    _need(raw_predictions, torch.float32, "raw_predictions")
    _need(normalized_cooccurrence_matrix, torch.float32, "normalized_cooccurrence_matrix")
    b, k = raw_predictions.shape
    assert tuple(normalized_cooccurrence_matrix.shape) == (k, k)
    out = torch.empty_like(raw_predictions)
    check(lib.lecb_cooc_adjust(_ptr(raw_predictions), _ptr(normalized_cooccurrence_matrix), _ptr(out), b, k, float(weight),
                               _stream()), "lecb_cooc_adjust")
    return out


def _block_fuse(data, sims, base, mode, threshold, weight):
    _need(data, torch.float32, "data")
    b, nb, k = data.shape
    sims_ld = 0
    if sims is not None:
        _need(sims, torch.float32, "sims_scores")
        assert tuple(sims.shape[:2]) == (b, nb)
        sims_ld = sims.shape[2]
    if base is not None:
        _need(base, torch.float32, "output")
        assert tuple(base.shape) == (b, k)
    out = torch.empty((b, k), device=data.device, dtype=torch.float32)
    check(lib.lecb_block_fuse(_ptr(data), _ptr(sims), sims_ld, _ptr(base), _ptr(out), b, nb, k, mode, float(threshold),
                              float(weight), _stream()), "lecb_block_fuse")
    return out


def aggregate_blocks(output, output_blocks, threshold=0.3, weight=1.4):
    """T:655-662: output_final = weight * s_ag + output, s_ag = max over windows if it exceeds `threshold` else min."""
    return _block_fuse(output_blocks, None, output, 0, threshold, weight)


def fuse(data, sims_scores, threshold=0.2):
    """gen_final_ans.py:18-36."""
    return _block_fuse(data, sims_scores, None, 1, threshold, 1.0)


def fuse6(data, sims_scores, threshold=0.2):
    """gen_final_ans.py:38-71."""
    return _block_fuse(data, sims_scores, None, 2, threshold, 1.0)
