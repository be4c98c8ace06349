"""BASELINE configs[3] on its own: the text-only prompt-tuning step (CLIP RN50 text tower + ASL, 77-token synthetic captions with
the real length distribution, 64 captions per GPU, prompt-gradient all-reduce, fused SGD) — the `extra.prompt_tuning` line of
bench.py, with switches for the A/B measurements.

    python tools/bench_train.py [--steps K] [--per-gpu-batch 64] [--no-graph] [--no-trim]
    python -m torch.distributed.run --nproc-per-node N tools/bench_train.py ...
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
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--per-gpu-batch", type=int, default=64)
    ap.add_argument("--no-graph", action="store_true", help="time the eager step (host launch overhead included)")
    ap.add_argument("--no-trim", action="store_true", help="run the caption tower on all 77 positions like the reference")
    args = ap.parse_args()
    cx = bench.Ctx()
    if args.no_trim:
        from lecb200 import dense_clip
        dense_clip.DenseCLIPB200.trim_caption_padding = False
    out = bench.bench_train(cx, args.per_gpu_batch, args.steps, args.warmup, use_graph=not args.no_graph)
    if cx.rank == 0:
        out["n_gpus"] = cx.world
        out["trim_caption_padding"] = not args.no_trim
        print(json.dumps(out), flush=True)
    if cx.world > 1:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
