"""Courtesy baseline: the reference's OWN PyTorch modules (clip/model.py `ModifiedResNet` + the dense head of
DenseCLIP.forward, T:405-472, written with the reference's tensor ops) timed in eager mode on the GPU — what a user of the
reference gets on this box without this repo.  Not the reference arm of bench.py (that one is the CPU path).

The reference tree does not exist on the GPU box: copy its model file next to this repo first (git-ignored, it travels
with gpurun):   mkdir -p baseline/_ref/clip && cp /root/reference/project/my_code/clip/model.py baseline/_ref/clip/

    python tools/bench_reference_gpu.py [--batch 64] [--dtype bf16|fp32|tf32] [--device cuda]
"""
import argparse
import importlib.util
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_reference_model_module():
    for base in (os.path.join(ROOT, "baseline", "_ref", "clip"), "/root/reference/project/my_code/clip"):
        path = os.path.join(base, "model.py")
        if os.path.exists(path):
            spec = importlib.util.spec_from_file_location("_ref_clip_model", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod, path
    raise SystemExit("reference clip/model.py not found: see the docstring for the copy command")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32", "tf32"])
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--res", type=int, default=448)
    args = ap.parse_args()
    M, path = load_reference_model_module()
    from oracle import synth
    arch = synth.RN101(args.res)
    dev = torch.device(args.device)
    clip = M.CLIP(*arch.ctor_args())
    clip.load_state_dict(synth.clip_state_dict(arch, 0), strict=False)
    visual = clip.visual.float().eval().to(dev)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = args.dtype == "tf32"
    k, d = 80, arch.embed_dim
    g = torch.Generator().manual_seed(0)
    t_pos, t_neg, t_evi = (F.normalize(torch.randn((k, d), generator=g), dim=-1).to(dev) for _ in range(3))
    images = torch.randn((args.batch, 3, args.res, args.res), generator=g).to(dev)
    ap_ = visual.attnpool

    @torch.no_grad()
    def step():
        # T:385-399 trunk, T:405-413 local + global features, T:441-472 head (evidence variant), all reference tensor ops
        x = images.type(visual.conv1.weight.dtype)
        for conv, bn in ((visual.conv1, visual.bn1), (visual.conv2, visual.bn2), (visual.conv3, visual.bn3)):
            x = visual.relu(bn(conv(x)))
        x = visual.avgpool(x)
        feat = visual.layer4(visual.layer3(visual.layer2(visual.layer1(x))))
        b, c, h, w = feat.shape
        tok = feat.reshape(b, c, h * w).permute(2, 0, 1)
        loc = F.linear(F.linear(tok, ap_.v_proj.weight, ap_.v_proj.bias), ap_.c_proj.weight, ap_.c_proj.bias)
        glob, _ = ap_(feat, if_pos=False)
        loc = loc / loc.norm(dim=-1, keepdim=True)
        glob = glob / glob.norm(dim=-1, keepdim=True)
        logits = 4.0 * glob @ t_pos.t()
        neg = loc @ t_neg.t()
        evi = loc @ t_evi.t()
        wta = F.softmax(50.0 * neg * (neg.max(dim=-1, keepdim=True)[0] + 1), dim=-1)
        prob = F.softmax(50.0 * evi, dim=0)
        return logits, (4.0 * neg * wta * prob).sum(0)

    ctx = torch.autocast(dev.type, dtype=torch.bfloat16) if args.dtype == "bf16" else torch.autocast(dev.type, enabled=False)
    with ctx:
        for _ in range(args.warmup):
            step()
        if dev.type == "cuda":
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
        else:
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            ms = (time.perf_counter() - t0) * 1e3 / args.steps
    print(json.dumps({"baseline": "reference PyTorch modules, eager", "source": path, "device": str(dev), "dtype": args.dtype,
                      "arch": f"RN101@{args.res}", "batch": args.batch, "ms_per_step": ms, "img_per_s": args.batch / (ms * 1e-3)}))


if __name__ == "__main__":
    main()
