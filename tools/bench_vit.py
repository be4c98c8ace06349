"""BASELINE configs[2] / [4] on their own: CLIP ViT-B/16 and ViT-L/14 dual-prompt inference at 448x448, batch-sharded across
the ranks with the packed logits all-gathered over NCCL — the `extra.vitb16` / `extra.vitl14` lines of bench.py.

    python tools/bench_vit.py --arch vitb16 [--batch 128] [--steps 10]
    python -m torch.distributed.run --nproc-per-node N tools/bench_vit.py --arch vitl14 --batch 128
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="vitb16", choices=["vitb16", "vitl14"])
    ap.add_argument("--batch", type=int, default=128, help="images per GPU per step")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    cx = bench.Ctx()
    out = bench.bench_vit(cx, args.arch, args.batch, args.steps, args.warmup)
    if cx.rank == 0:
        out["n_gpus"] = cx.world
        print(json.dumps(out), flush=True)
    if cx.world > 1:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
