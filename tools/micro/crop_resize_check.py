"""Checker for tools/micro/crop_resize_probe.cu: sliding windows of a random image (lecb200.windows) -> probe binary ->
compare with the oracle (numpy crop + oracle/pil_resize.py, bit-exact against Pillow).

    python tools/micro/crop_resize_check.py [--emulate] [--scales 2 3] [--size 448]

--emulate replaces the GPU binary by a numpy transcription of the two kernels (same plan arrays, same index arithmetic): it
validates the composition windows -> reflect/crop rows -> horizontal taps -> uint8 -> vertical taps on the CPU, which is
what could be checked in round 1 (no GPU minutes left); without it the binary must exist (see the .cu header) and a GPU."""
import argparse
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
MEAN, STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)


def emulate(img, wins, S):
    from lecb200 import windows as WN
    H = img.shape[0]
    out = np.empty((len(wins), S, S, 3), dtype=np.uint8)
    for i, w in enumerate(wins):
        hb, hk = WN.resize_plan(w.width, S)
        vb, vk = WN.resize_plan(w.height, S)
        rows = np.array([WN.padded_row_source(w.top + y, H, w.pad_top, w.pad_bottom) for y in range(w.height)])
        acc = np.full((w.height, S, 3), 1 << 21, dtype=np.int64)                    # resize_h_kernel
        for t in range(hk.shape[1]):
            live = (t < hb[:, 1])[None, :, None]
            cols = w.left + np.minimum(hb[:, 0] + t, w.width - 1)
            acc += np.where(live, hk[:, t][None, :, None].astype(np.int64) * img[rows][:, cols].astype(np.int64), 0)
        tmp = np.clip(acc >> 22, 0, 255).astype(np.uint8)
        acc = np.full((S, S, 3), 1 << 21, dtype=np.int64)                           # resize_v_kernel
        for t in range(vk.shape[1]):
            live = (t < vb[:, 1])[:, None, None]
            src = np.minimum(vb[:, 0] + t, w.height - 1)
            acc += np.where(live, vk[:, t][:, None, None].astype(np.int64) * tmp[src].astype(np.int64), 0)
        out[i] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
    f32 = ((out.astype(np.float32) / np.float32(255.0) - np.asarray(MEAN, np.float32)) / np.asarray(STD, np.float32)).transpose(0, 3, 1, 2)
    return out, np.ascontiguousarray(f32)


def run_probe(img, wins, S):
    exe = os.path.join(ROOT, "tools", "micro", "crop_resize_probe")
    if not os.path.exists(exe):
        raise SystemExit("build tools/micro/crop_resize_probe first (command in the .cu header)")
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("4i", img.shape[0], img.shape[1], S, len(wins)))
            for w in wins:
                f.write(struct.pack("6i", w.top, w.left, w.height, w.width, w.pad_top, w.pad_bottom))
            f.write(img.tobytes())
        print(subprocess.run([exe, fin, fout], check=True, capture_output=True, text=True).stdout.strip())
        raw = open(fout, "rb").read()
    n = len(wins) * S * S * 3
    return (np.frombuffer(raw[:n], dtype=np.uint8).reshape(len(wins), S, S, 3),
            np.frombuffer(raw[n:], dtype=np.float32).reshape(len(wins), 3, S, S))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--emulate", action="store_true")
    ap.add_argument("--scales", type=int, nargs="+", default=[2, 3])
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--image", type=int, nargs=2, default=[375, 500])
    args = ap.parse_args()
    from lecb200 import windows as WN
    from oracle import pil_resize as PR
    rng = np.random.default_rng(3)
    h, w = args.image
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    wins = [x for s in args.scales for x in WN.sliding_windows(h, w, s)]
    got_u8, got_f32 = (emulate if args.emulate else run_probe)(img, wins, args.size)
    bad = 0
    for i, win in enumerate(wins):
        rows, cols = WN.source_rows_cols(win, h, w)
        crop = np.ascontiguousarray(img[rows][:, cols])
        want = PR.resize_u8(crop, args.size, args.size, "bicubic")
        if not np.array_equal(got_u8[i], want):
            bad += 1
            print(f"window {i} {win}: {int((got_u8[i] != want).sum())} bytes differ")
        elif np.abs(got_f32[i] - PR.test_transform(crop, (args.size, args.size), MEAN, STD)).max() > 1e-6:
            bad += 1
            print(f"window {i}: float output off")
    print(f"{len(wins)} windows, {bad} mismatching -> {'OK' if bad == 0 else 'FAIL'} ({'numpy emulation' if args.emulate else 'GPU probe'})")
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
