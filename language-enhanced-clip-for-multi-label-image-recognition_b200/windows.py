"""Sliding-window geometry of the test-time multi-window pipeline, as pure index arithmetic.

Reference: `DatasetWrapperWithBlock._transform_image`, dassl/data/data_manager.py:348-492 — for every scale `block_size` of
`multi_scale` it cuts, in this order, (1) a (2s x 2s) grid of (h/s x w/s) windows out of the reflect-padded image (see
grid_padding for what the padding really is), (2) 1x2 and 2x1 windows, (3) 1x1.5 and 1.5x1 windows, (4) for s >= 3, 2x3 and 3x2 windows, groups 2-4 clipped
at the image border and dropped when empty; every window is then resized to the network input by the dataset transform.
This module reproduces the rectangles (bit-exact integers, pinned to the reference's own code in
tests/test_windows.py / tests/golden/windows.npz); the crop + resize itself still runs in the reference's PIL pipeline —
a GPU kernel for it is the next §8f-1 step and will consume exactly these rectangles.  Pure host code: no kernels."""
from __future__ import annotations

from typing import List, NamedTuple, Sequence, Tuple


class Window(NamedTuple):
    top: int            # first row; for group-1 windows in the coordinates of the row-padded image (see grid_padding)
    left: int
    height: int
    width: int
    pad_top: int        # group 1: rows of reflect padding above / below the image that `top` counts through; a negative
    pad_bottom: int     # value crops instead (torchvision's F.pad semantics); (0, 0) for groups 2-4


def grid_padding(h: int, w: int, s: int) -> Tuple[int, int, int, int]:
    """(stride_h, stride_w, pad_top, pad_bottom) of the 2s x 2s grid of group 1 (data_manager.py:383-388).

    Reference quirk, reproduced: the code computes a bottom and a right padding and calls torchvision's
    `F.pad(img, (0, padding_w, 0, padding_h))` — whose 4-tuple means (left, TOP, right, bottom).  So `padding_w` rows are
    reflected ABOVE the image, `padding_h` rows below it, and no column is padded: the windows of the last grid column are
    simply clipped at the right border by the tensor slice (and come out narrower).  Either value can be negative, which
    torchvision turns into a crop of that many rows before the (positive) padding is applied."""
    slide = 2 * s
    bh, bw = h // s, w // s
    sh = ((s - 1) * bh) // (slide - 1) + 1
    sw = ((s - 1) * bw) // (slide - 1) + 1
    padding_h = sh * (slide - 1) - (s - 1) * bh - h % s
    padding_w = sw * (slide - 1) - (s - 1) * bw - w % s
    return sh, sw, padding_w, padding_h


def padded_rows(h: int, pad_top: int, pad_bottom: int) -> int:
    """Row count of the image after torchvision's F.pad with (top, bottom) = (pad_top, pad_bottom)."""
    return h + pad_top + pad_bottom


def padded_row_source(p: int, h: int, pad_top: int, pad_bottom: int = 0) -> int:
    """Source row of row p of the padded image: negative paddings crop first, positive ones then reflect about the
    edges of the cropped image (torch 'reflect': no edge repeat)."""
    crop_top, crop_bottom = max(-pad_top, 0), max(-pad_bottom, 0)
    he = h - crop_top - crop_bottom
    q = p - max(pad_top, 0)
    if q < 0:
        q = -q
    elif q >= he:
        q = 2 * (he - 1) - q
    return crop_top + q


def _clipped_group(h, w, s, bh, bw, nh, nw) -> List[Window]:
    sh = ((s - 1) * bh) // (nh - 1) + 1
    sw = ((s - 1) * bw) // (nw - 1) + 1
    out = []
    for i in range(nh):
        for j in range(nw):
            ch, cw = min(bh, h - i * sh), min(bw, w - j * sw)
            if ch <= 0 or cw <= 0:
                continue
            out.append(Window(i * sh, j * sw, ch, cw, 0, 0))
    return out


def sliding_windows(h: int, w: int, s: int) -> List[Window]:
    """All windows of one scale `s` (= block_size) of an h x w image, in the reference's order."""
    sh, sw, pad_top, pad_bottom = grid_padding(h, w, s)
    he = h - max(-pad_top, 0) - max(-pad_bottom, 0)
    if he <= 0 or max(pad_top, 0) >= he or max(pad_bottom, 0) >= he:
        raise ValueError(f"sliding_windows: padding ({pad_top}, {pad_bottom}) invalid for a {h} x {w} image at scale {s}")
    bh, bw = h // s, w // s
    hp = padded_rows(h, pad_top, pad_bottom)
    wins = []
    for i in range(2 * s):
        for j in range(2 * s):
            ch, cw = min(bh, hp - i * sh), min(bw, w - j * sw)          # the slice clips at the padded / right border
            if ch <= 0 or cw <= 0:
                raise ValueError(f"sliding_windows: empty grid window ({i}, {j}) for a {h} x {w} image at scale {s}")
            wins.append(Window(i * sh, j * sw, ch, cw, pad_top, pad_bottom))
    wins += _clipped_group(h, w, s, h // s, w * 2 // s, 2 * s, s)                       # 1 x 2
    wins += _clipped_group(h, w, s, h * 2 // s, w // s, s, 2 * s)                       # 2 x 1
    wins += _clipped_group(h, w, s, h // s, w * 3 // (2 * s), 2 * s, 2 * s * 2 // 3)    # 1 x 1.5
    wins += _clipped_group(h, w, s, h * 3 // (2 * s), w // s, 2 * s * 2 // 3, 2 * s)    # 1.5 x 1
    if s >= 3:
        wins += _clipped_group(h, w, s, h * 2 // s, w * 3 // s, 2 * s // 2, 2 * s // 3)  # 2 x 3
        wins += _clipped_group(h, w, s, h * 3 // s, w * 2 // s, 2 * s // 3, 2 * s // 2)  # 3 x 2
    return wins


def is_grid(win: Window, index: int, s: int) -> bool:
    """The first 4 s^2 windows of a scale are the (padded) grid of group 1."""
    return index < 4 * s * s


def windows_for_scales(h: int, w: int, multi_scale: Sequence[int] = (2, 3, 4, 5)) -> List[List[Window]]:
    """One list per scale, like the `img_blocks` list of the reference (default scales: data_manager.py:313)."""
    return [sliding_windows(h, w, s) for s in multi_scale]


def source_rows_cols(win: Window, h: int, w: int):
    """Source row / column indices of a window's pixels in the unpadded image (reflection resolved)."""
    rows = [padded_row_source(win.top + y, h, win.pad_top, win.pad_bottom) for y in range(win.height)]
    cols = [win.left + x for x in range(win.width)]
    return rows, cols


def resize_plan(in_size: int, out_size: int, filt: str = "bicubic"):
    """Pillow-compatible taps for resizing one axis (lecb_resize_plan, host code of liblecb.so) ->
    (bounds int32 [out,2] = (first source index, tap count), coeffs int32 [out, ksize] in 22-bit fixed point)."""
    import ctypes

    import numpy as np

    from . import _lib
    f = {"bilinear": 0, "bicubic": 1}[filt]
    ksize = _lib.lib.lecb_resize_ksize(in_size, out_size, f)
    _lib.check(min(ksize, 0), "lecb_resize_ksize")
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    _lib.check(_lib.lib.lecb_resize_plan(in_size, out_size, f, bounds.ctypes.data_as(ctypes.c_void_p),
                                         coeffs.ctypes.data_as(ctypes.c_void_p), ksize), "lecb_resize_plan")
    return bounds, coeffs


_PLAN_CACHE = {}


def crop_resize(img, wins: Sequence[Window], size: int, filt: str = "bicubic", want_u8: bool = True, want_f32: bool = False,
                mean=None, std=None):
    """Crop + Pillow-compatible resize (+ ToTensor + Normalize) of `wins` of ONE decoded image on the GPU
    (lecb_crop_resize_u8; reference: data_manager.py:348-492 + transforms.py:379-411, about a hundred PIL calls per image).

    img uint8 CUDA tensor [H,W,3].  -> (u8 NHWC [n,size,size,3] | None, f32 NCHW [n,3,size,size] | None); the uint8 batch
    is byte-identical to `PIL.Image.resize` of each crop and feeds `DenseCLIPB200(image_u8, if_test=True)` directly."""
    import ctypes

    import numpy as np
    import torch

    from . import _lib, ops
    if not (img.is_cuda and img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3 and img.is_contiguous()):
        raise _lib.LecbError("crop_resize: img must be a contiguous uint8 CUDA tensor [H,W,3]")
    h, w = int(img.shape[0]), int(img.shape[1])
    n = len(wins)
    f = {"bilinear": 0, "bicubic": 1}[filt]
    # the plan depends on the image SIZE and the window list only: datasets have a handful of image sizes, so the uploaded
    # blob and the scratch buffer are kept per (size, windows) and an image costs just the two kernel launches
    key = (h, w, size, f, str(img.device), tuple(wins))
    hit = _PLAN_CACHE.get(key)
    if hit is None:
        rec = np.ascontiguousarray(np.asarray([[x.top, x.left, x.height, x.width, x.pad_top, x.pad_bottom] for x in wins], dtype=np.int32))
        plan_ints, tmp_bytes = ctypes.c_longlong(0), ctypes.c_longlong(0)
        wp = rec.ctypes.data_as(ctypes.c_void_p)
        _lib.check(_lib.lib.lecb_window_plan_size(wp, n, h, w, size, f, ctypes.byref(plan_ints), ctypes.byref(tmp_bytes)),
                   "lecb_window_plan_size")
        plan = torch.empty((plan_ints.value,), dtype=torch.int32)
        _lib.check(_lib.lib.lecb_window_plan(wp, n, h, w, size, f, plan.data_ptr(), plan_ints.value), "lecb_window_plan")
        if len(_PLAN_CACHE) >= 64:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        hit = _PLAN_CACHE[key] = (plan.to(img.device), torch.empty((max(tmp_bytes.value, 1),), device=img.device, dtype=torch.uint8))
    plan_d, tmp = hit
    out_u8 = torch.empty((n, size, size, 3), device=img.device, dtype=torch.uint8) if want_u8 else None
    out_f32 = torch.empty((n, 3, size, size), device=img.device, dtype=torch.float32) if want_f32 else None
    m, s = ops._f3(mean or ops.CLIP_PIXEL_MEAN), ops._f3(std or ops.CLIP_PIXEL_STD)
    _lib.check(_lib.lib.lecb_crop_resize_u8(img.data_ptr(), h, w, plan_d.data_ptr(), n, size, tmp.data_ptr(), ops._ptr(out_u8),
                                            ops._ptr(out_f32), m, s, ops._stream()), "lecb_crop_resize_u8")
    return out_u8, out_f32


def whole_image(h: int, w: int) -> Window:
    """The un-windowed test image as a Window (the first input of every sample: `Resize(INPUT.SIZE)` of the full image)."""
    return Window(0, 0, h, w, 0, 0)
