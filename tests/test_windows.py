"""Sliding-window geometry against the reference's own `_transform_image` (data_manager.py:348-492).

tests/golden/windows.npz was produced by oracle/make_windows_golden.py: the reference code, unmodified, run on images whose
pixels encode their coordinates, so each block it returns reveals its source rows and columns.  Index arithmetic: the
comparison is exact.  The live test repeats it per pixel where /root/reference exists."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import ref_extract as RX

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _module():
    spec = importlib.util.spec_from_file_location(
        "_lecb200_windows", os.path.join(ROOT, "language-enhanced-clip-for-multi-label-image-recognition_b200", "windows.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


W = _module()
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "windows.npz"))


def _summary(h, w, s):
    rows = []
    for win in W.sliding_windows(h, w, s):
        r, c = W.source_rows_cols(win, h, w)
        rows.append([r[0], r[-1], len(r), c[0], c[-1], len(c)])
    return np.array(rows, dtype=np.int32)


@pytest.mark.parametrize("key", sorted(GOLD.files))
def test_windows_match_reference_golden(key):
    size, s = key.split("_s")
    h, w = (int(v) for v in size.split("x"))
    got, want = _summary(h, w, int(s)), GOLD[key]
    assert got.shape == want.shape, f"{key}: {got.shape[0]} windows, reference has {want.shape[0]}"
    assert np.array_equal(got, want), f"{key}: first mismatch at window {int(np.argwhere((got != want).any(1))[0, 0])}"


def test_window_counts_and_padding():
    # the four default scales: 4 s^2 grid windows + the clipped groups (empty windows of groups 2-4 are dropped)
    counts = [len(v) for v in W.windows_for_scales(224, 224)]
    assert counts == [GOLD[f"224x224_s{s}"].shape[0] for s in (2, 3, 4, 5)] and counts[0] == 40
    for h, w in ((448, 448), (375, 500), (97, 131)):
        for s in (2, 3, 4, 5):
            sh, sw, pad_top, pad_bottom = W.grid_padding(h, w, s)
            wins = W.sliding_windows(h, w, s)
            for k, x in enumerate(wins):
                assert x.height > 0 and x.width > 0 and x.left + x.width <= w
                if W.is_grid(x, k, s):
                    assert (x.pad_top, x.pad_bottom) == (pad_top, pad_bottom) and x.top + x.height <= h + pad_top + pad_bottom
                else:
                    assert (x.pad_top, x.pad_bottom) == (0, 0) and x.top + x.height <= h
    # reference quirk (F.pad's 4-tuple is left, top, right, bottom): rows are reflected above the image too
    assert [W.padded_row_source(p, 5, 2, 2) for p in range(9)] == [2, 1, 0, 1, 2, 3, 4, 3, 2]
    assert [W.padded_row_source(p, 6, 1, -2) for p in range(5)] == [1, 0, 1, 2, 3]          # negative padding crops
    assert W.grid_padding(375, 500, 4) == (40, 54, 3, -2)
    with pytest.raises(ValueError):
        W.sliding_windows(3, 400, 5)


@pytest.mark.skipif(not RX.available(), reason="/root/reference not present")
@pytest.mark.parametrize("h,w", [(240, 320), (101, 77)])
def test_every_pixel_live(h, w):
    from oracle.make_windows_golden import SCALES, reference_windows
    for s, ref in zip(SCALES, reference_windows(h, w)):
        ours = W.sliding_windows(h, w, s)
        assert len(ours) == len(ref)
        for win, (rows, cols) in zip(ours, ref):
            r, c = W.source_rows_cols(win, h, w)
            assert np.array_equal(np.asarray(r), rows) and np.array_equal(np.asarray(c), cols)
