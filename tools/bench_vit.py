"""BASELINE configs[2] / [4]: CLIP ViT-B/16 and ViT-L/14 dual-prompt inference at 448x448, batch-sharded across the
ranks with the packed logits all-gathered over NCCL (secondary measurement; bench.py carries the headline).

    python tools/bench_vit.py --arch vitb16 [--batch 128] [--steps 10] [--profile]
    python -m torch.distributed.run --nproc-per-node N tools/bench_vit.py --arch vitl14 --batch 128
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="vitb16", choices=["vitb16", "vitl14"])
    ap.add_argument("--batch", type=int, default=128, help="images per GPU per step")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import lecb200
    from bench import load_tokens, make_cfg
    from lecb200 import synth
    from lecb200.clip_model import CLIPParams
    from lecb200.dense_clip import DenseCLIPB200
    from lecb200.dist import all_gather_logits
    from lecb200.prof import KernelTimer

    arch = synth.VITB16(448) if args.arch == "vitb16" else synth.VITL14(448)
    toks, n_ctx, names = load_tokens()
    clip = CLIPParams(*arch.ctor_args())
    clip.load_state_dict(synth.clip_state_dict(arch, 0), strict=False)
    clip = clip.float().to(dev).eval()
    model = DenseCLIPB200(make_cfg(448, n_ctx, True), names, clip, tokenized_prompts=toks).to(dev)
    images = torch.randn((args.batch, 3, 448, 448), device=dev)

    def step():
        out = model(images, if_test=True)
        return all_gather_logits(out[0], out[1])

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = lecb200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    gf_img = 156.99 + 1.5 if args.arch == "vitb16" else 723.59 + 3.0          # SURVEY §8(d), algorithmic GF per image
    if rank == 0:
        line = {"metric": "multi_label_images_per_sec_80cls_448px", "arch": args.arch, "value": world * args.batch / (ms * 1e-3),
                "unit": "img/s", "n_gpus": world, "per_gpu_batch": args.batch, "ms_per_step": ms,
                "tflops_per_gpu_algorithmic": gf_img * args.batch / ms, "gpu_launches_per_step": (lecb200.launch_count() - n0) // args.steps,
                "finite": bool(torch.isfinite(res[0]).all() and torch.isfinite(res[1]).all()),
                "parity": "global feature pinned to the reference VisionTransformer; dense head vs repo oracle (reference has no ViT dense path)"}
        print(json.dumps(line), flush=True)
        if args.profile:
            with KernelTimer() as kt:
                for _ in range(2):
                    model(images, if_test=True)
            rows = kt.detail(2)
            for r in rows[:14]:
                print(r)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
