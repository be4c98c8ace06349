"""TEST INFRASTRUCTURE (oracle): numpy restatement of Pillow's `Image.resize` for 8-bit images, the arithmetic behind the
reference's test-time transform (`Resize(..., interpolation=bicubic)` on PIL images, dassl/data/transforms/transforms.py:
384-394, applied to the whole image and to every sliding window, data_manager.py:348-492).

Third-party algorithm: Pillow (python-pillow/Pillow, src/libImaging/Resample.c — `precompute_coeffs`,
`normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`); the reference pins no version
(its Dockerfile installs whatever pip resolves), this container has Pillow 12.2.0, against which tests/test_pil_resize.py
checks the restatement BIT FOR BIT.  The algorithm: a separable convolution — horizontal pass first, rounded to uint8, then
the vertical pass — whose per-output-pixel taps are the filter sampled at the source pixel centres inside
`support * max(scale, 1)` of the output pixel's centre (this widening is the antialiasing), normalised to sum 1 in double
precision, converted to 22-bit fixed point with round-half-away-from-zero, accumulated in 32-bit integers from a
half-unit start value, shifted down and clamped to [0, 255].  Nothing here is on the product path."""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear(x: float) -> float:
    if x < 0.0:
        x = -x
    return 1.0 - x if x < 1.0 else 0.0


FILTERS = {"bicubic": (_bicubic, 2.0), "bilinear": (_bilinear, 1.0)}


def precompute_coeffs(in_size: int, in0: float, in1: float, out_size: int, filt: str):
    """-> (bounds int [out,2] = (first source index, tap count), coefficients int32 [out, ksize] in 22-bit fixed point)."""
    fn, support0 = FILTERS[filt]
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            w = fn((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    # normalize_coeffs_8bpc: (int)(+-0.5 + k * 2^22)
    scaled = kk * (1 << PRECISION_BITS)
    fixed = np.where(kk < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int64)
    return bounds, fixed


def _resample_axis0(img: np.ndarray, out_size: int, filt: str) -> np.ndarray:
    """Resample axis 0 of a uint8 array [n, ...]."""
    n = img.shape[0]
    bounds, kk = precompute_coeffs(n, 0.0, float(n), out_size, filt)
    src = img.astype(np.int64)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        x0, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[x0 + x] * kk[xx, x]
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_u8(img: np.ndarray, out_h: int, out_w: int, filt: str = "bicubic") -> np.ndarray:
    """img uint8 [H, W, C] -> uint8 [out_h, out_w, C], as `PIL.Image.resize((out_w, out_h), resample=<filt>)`.
    Like ImagingResample, a pass whose size does not change is skipped, and the horizontal pass runs first."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, _ = img.shape
    cur = img
    if out_w != w:
        cur = np.ascontiguousarray(np.swapaxes(_resample_axis0(np.ascontiguousarray(np.swapaxes(cur, 0, 1)), out_w, filt), 0, 1))
    if out_h != h:
        cur = _resample_axis0(cur, out_h, filt)
    return cur


def test_transform(img: np.ndarray, size, mean, std, filt: str = "bicubic") -> np.ndarray:
    """`Resize(size) -> ToTensor -> Normalize` (transforms.py:392-402, no center crop): uint8 [H,W,3] -> float32 [3,h,w]."""
    out = resize_u8(img, size[0], size[1], filt).astype(np.float32) / np.float32(255.0)
    out = (out - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
    return np.ascontiguousarray(out.transpose(2, 0, 1))
