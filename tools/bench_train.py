"""BASELINE configs[3]: text-only prompt-tuning step (CLIP text tower + ASL), 77-token synthetic captions,
64 captions per GPU (global 512 at 8 GPUs), prompt-gradient all-reduce, SGD step.  Secondary measurement
(bench.py carries the headline images/sec metric).

    python tools/bench_train.py [--steps K] [--per-gpu-batch 64] [--evidence]
    python -m torch.distributed.run --nproc-per-node N tools/bench_train.py ...
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--per-gpu-batch", type=int, default=64)
    ap.add_argument("--evidence", action="store_true")
    ap.add_argument("--loss", default="asl", choices=["asl", "ranking"])
    ap.add_argument("--shard-prompts", action="store_true",
                    help="split the prompt sequences over the ranks instead of replicating them (DenseCLIPB200.shard_prompt_branch)")
    ap.add_argument("--graph", action="store_true", help="capture forward+loss+backward+allreduce+SGD in one CUDA graph")
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import lecb200
    from bench import load_tokens, make_cfg
    from lecb200 import losses, synth
    from lecb200.clip_model import CLIPParams
    from lecb200.dense_clip import DenseCLIPB200
    from lecb200.dist import allreduce_mean_grads, broadcast_params

    arch = synth.RN50(224)
    toks, n_ctx, names = load_tokens()
    clip = CLIPParams(*arch.ctor_args())
    clip.load_state_dict(synth.clip_state_dict(arch, 0), strict=False)
    clip = clip.float().to(dev).eval()
    model = DenseCLIPB200(make_cfg(224, n_ctx, args.evidence), names, clip, tokenized_prompts=toks).to(dev)
    for n_, p in model.named_parameters():
        p.requires_grad_("prompt_learner." in n_ and "prompt_learner_m" not in n_)
    model.shard_prompt_branch = bool(args.shard_prompts)
    params = [p for p in model.prompt_learner.parameters()]
    broadcast_params(params)                 # DDP's construction-time broadcast (T:786-787)
    opt = torch.optim.SGD(params, lr=0.002, momentum=0.9)
    b = args.per_gpu_batch
    caps = synth.captions(b, 100 + rank, vocab=arch.vocab_size).to(dev)
    y = synth.labels(b, len(names), 100 + rank).to(dev)
    loss_fn = losses.ASL_loss if args.loss == "asl" else (lambda o, t: losses.ranking_loss(o, t, scale_=1.0, margin_=1))

    def step():
        out = model(None, caps)
        loss = loss_fn(out[0], y) + loss_fn(out[1], y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        allreduce_mean_grads(params)
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    if args.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            static_loss = step()
        eager_step = step

        def step():
            graph.replay()
            return static_loss
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = lecb200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n_seq = 3 if args.evidence else 2
    # 6.04 GF per 77-token sequence forward; prompt branch fwd + dgrad-only bwd ~ 2x its forward
    tf = (b * 6.04 + n_seq * len(names) * 5.96 * 3.0) / 1e3
    if rank == 0:
        print(json.dumps({"metric": "prompt_tuning_captions_per_sec", "value": world * b / (ms * 1e-3), "unit": "captions/s",
                          "n_gpus": world, "ms_per_step": ms, "per_gpu_batch": b, "global_batch": world * b,
                          "prompt_sequences": n_seq * len(names), "loss": args.loss, "cuda_graph": bool(args.graph), "shard_prompts": bool(args.shard_prompts), "final_loss": float(loss),
                          "approx_tflops_per_rank": tf / (ms * 1e-3), "gpu_launches_per_step": (lecb200.launch_count() - n0) // args.steps,
                          "config": "BASELINE configs[3] shape: RN50 text tower, 77 tokens, real caption-length distribution"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
