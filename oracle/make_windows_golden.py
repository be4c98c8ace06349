"""Golden window rectangles for tests/test_windows.py, produced by running the reference's own
`DatasetWrapperWithBlock._transform_image` (AST-extracted, unmodified) on coordinate-coded PIL images.

    python -m oracle.make_windows_golden          # writes tests/golden/windows.npz (needs /root/reference)

Every pixel of the probe image encodes its own (row, col) in its three uint8 channels, the dataset transform is replaced by
`to_tensor`, so each returned block tells where it was cut from — including the reflect-padded rows / columns."""
import os
import types

import numpy as np
import torch

from . import ref_extract as RX

SIZES = [(224, 224), (480, 640), (375, 500), (333, 500), (500, 281), (97, 131)]
SCALES = [2, 3, 4, 5]


def coded_image(h, w):
    from PIL import Image
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    rgb = np.stack([xx & 255, yy & 255, (xx >> 8) | ((yy >> 8) << 4)], -1).astype(np.uint8)
    return Image.fromarray(rgb)


def decode(block):
    """block float [3,bh,bw] in [0,1] -> (rows [bh], cols [bw]) source indices; asserts the block is a separable crop."""
    u = (block * 255.0).round().to(torch.int64)
    cols = (u[0] | ((u[2] & 15) << 8))
    rows = (u[1] | ((u[2] >> 4) << 8))
    assert (rows == rows[:, :1]).all() and (cols == cols[:1, :]).all(), "not an axis-aligned crop"
    return rows[:, 0].numpy(), cols[0, :].numpy()


def reference_windows(h, w, scales=SCALES):
    import torchvision.transforms.functional as F
    fn = RX.window_transform()
    me = types.SimpleNamespace(k_tfm=1, multi_scale=list(scales))

    class _Keep(list):                       # the reference stacks the blocks of a scale: keep them ragged instead
        pass

    blocks_per_scale = []
    real_stack = torch.stack
    try:
        torch.stack = lambda seq, *a, **k: _Keep(seq)
        _, img_blocks = fn(me, lambda pil: F.to_tensor(pil), coded_image(h, w))
    finally:
        torch.stack = real_stack
    for blocks in img_blocks:
        blocks_per_scale.append([decode(b) for b in blocks])
    return blocks_per_scale


def main():
    out = {}
    for h, w in SIZES:
        per_scale = reference_windows(h, w)
        for s, wins in zip(SCALES, per_scale):
            # first / last source row and column + extent of every window, in the reference's order
            arr = np.array([[r[0], r[-1], len(r), c[0], c[-1], len(c)] for r, c in wins], dtype=np.int32)
            out[f"{h}x{w}_s{s}"] = arr
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "windows.npz")
    np.savez_compressed(path, **out)
    print(f"{len(out)} (size, scale) cases -> {path}")


if __name__ == "__main__":
    main()
