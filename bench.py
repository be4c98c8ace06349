#!/usr/bin/env python
"""bench.py — headline metric of the dual-prompt scoring path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): multi-label images/sec, 80 classes, 448x448.  Workload at every N = BASELINE
configs[1] per GPU: CLIP RN101 dual-prompt inference (pos/neg/evidence prompts, 80 COCO classes), synthetic
448x448 images, batch 256 per GPU, bf16 operands / fp32 accumulation, random-init weights (weak scaling:
per-GPU batch fixed, the packed logits all-gathered over NCCL inside the timed region).

One JSON line on rank 0:
  value       img/s with the input batch resident in HBM (CUDA events, max over ranks)
  e2e         img/s through the public API with pinned HOST images: H2D copy of every step's input and D2H
              read of its logits inside the timed region (double-buffered copy stream).  The host batch is RAW
              uint8 NHWC pixels (what an image decoder produces; ToTensor + Normalize run inside the stem kernel,
              bit-identical to the reference's float tensor); `e2e_fp32` is the same loop fed with the
              reference-format normalised float NCHW batch (4x the bytes over PCIe)
  gpu_reference  the reference's OWN PyTorch modules (staged copy of clip/model.py + the head of T:405-472), eager,
              on this GPU: what a user of the reference gets on this box (bf16 autocast and TF32)
  extra       BASELINE configs 3 / 4 / 5 at the same N: ViT-B/16, ViT-L/14 inference and the prompt-tuning step
  roofline    dominant kernel = the tcgen05 GEMM / implicit-GEMM conv kernel: algorithmic FLOPs per step /
              its summed CUDA-event duration, vs the measured sustained bf16 peak
  cpu_baseline / --impl reference: the reference's own `DenseCLIP.forward(image, if_test=True)` (kind "reference":
              its source files are staged into the git-ignored baseline/_ref/ by build() and executed unmodified
              through oracle/ref_extract.py) on the box's host cores; falls back to the restatement
              (oracle/restatement.py, kind "port") only when the staged tree is missing.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOAD = "CLIP RN101 dual-prompt (pos/neg/evidence) inference, 80 classes, 448x448 synthetic images, batch 256 per GPU"
METRIC = "multi_label_images_per_sec_80cls_448px"
UNIT = "img/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def load_tokens():
    tk = np.load(os.path.join(ROOT, "tests", "golden", "prompt_tokens_coco80.npz"), allow_pickle=False)
    return torch.from_numpy(tk["tokens"]), int(tk["n_ctx"]), [str(s) for s in tk["classnames"]]


class Cfg(dict):
    __getattr__ = dict.__getitem__


def product_config(world, B, K):
    return {"workload": WORKLOAD, "global_batch": world * B, "per_gpu_batch": B, "classes": K,
            "parallelism": f"dp{world}", "l2_policy": "input batch (616 MB/GPU) larger than L2, no flush",
            "collective": "all_gather_into_tensor of packed [B,2K] fp32 logits per step" if world > 1 else "none"}


def make_cfg(res, n_ctx, use_evidence):
    return Cfg(TRAINER=Cfg(Caption=Cfg(N_CTX=n_ctx, CTX_INIT="", CSC=False, CLASS_TOKEN_POSITION="end",
                                       use_evidence=use_evidence, PREC="fp32")),
               INPUT=Cfg(SIZE=(res, res)),
               TRAIN=Cfg(IF_LEARN_SCALE=False, IF_LEARN_spatial_SCALE=False, spatial_SCALE_text=50.0,
                         spatial_SCALE_image=50.0, ema=False, momentum=0.995, LOSSFUNC="double_ranking"))


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.tmp = index, None, None

    def start(self):
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples whose timestamp falls inside [t_begin, t_end] (time.time() values)."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.tmp.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t_begin is not None and not (t_begin - 0.05 <= ts <= t_end + 0.05):
                    continue
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for n, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        return out


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle restatement of the reference image path (text features cached like T:421-439)
# --------------------------------------------------------------------------------------------------
def cpu_reference_model(arch):
    """The reference's own DenseCLIP (AST-extracted from the staged baseline/_ref tree, executed unmodified) holding the
    synthetic RN101 weights and the 80 COCO prompts; None when the staged tree is not there."""
    from oracle import stage_reference
    root = stage_reference.staged_root()
    if root is None:
        return None
    os.environ["LECB_REFERENCE_ROOT"] = root
    import importlib
    from oracle import ref_extract as RX
    RX = importlib.reload(RX)
    from oracle import synth
    # T:445 needs the module-level caption bank; the product arm runs without one (SURVEY 8c deviation 1: the 1024 literal
    # of T:447), so the reference gets a 16-row bank whose cost is nil
    bank = synth.caption_bank(16, arch.embed_dim, 0)
    ns = RX.trainer_classes(bank, arch.embed_dim)
    clip_model = RX.build_reference_clip(arch, synth.clip_state_dict(arch, 0))
    cfg = RX.make_cfg(arch.image_resolution, n_ctx=16, use_evidence=True)
    model = ns["DenseCLIP"](cfg, RX.coco_classnames(), clip_model).eval()
    return model


def cpu_reference_rate(arch, steps, warmup, budget_s, threads=None):
    from oracle import synth
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    ref_model = None
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints while it builds its prompts
            ref_model = cpu_reference_model(arch)
    except Exception as e:  # pragma: no cover
        print(f"[bench] reference classes unavailable ({type(e).__name__}: {e}); timing the port", file=sys.stderr)
    if ref_model is not None:
        def step(img):
            with torch.no_grad():
                return ref_model(img, if_test=True)

        probe = synth.images(1, arch.image_resolution, 99)
        step(probe)                                   # first call encodes and caches the 240 prompt features (T:421-439)
        t0 = time.perf_counter()
        step(probe)
        t_img = time.perf_counter() - t0
        total_steps = steps + warmup
        per_step = max(1, min(32, int(budget_s / max(t_img, 1e-3) / max(total_steps, 1))))
        imgs = synth.images(per_step, arch.image_resolution, 100)
        for _ in range(warmup):
            step(imgs)
        t0 = time.perf_counter()
        for _ in range(steps):
            step(imgs)
        dt = time.perf_counter() - t0
        return {"value": per_step * steps / dt, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"{steps} steps x {per_step} synthetic 448x448 images through the reference's own "
                          f"DenseCLIP.forward(image, if_test=True) (Caption_distill_double.py:402-472, staged unmodified under "
                          f"baseline/_ref, RN101 weights of the product arm, fp32, torch CPU, prompt features cached as T:421-439); "
                          f"the step is capped at {per_step} images because the CPU path runs at ~10 img/s: the workload's "
                          f"256-image step would take ~25 s and the K+W steps of one run more than ten minutes",
                "ms_per_step": 1e3 * dt / steps, "images_per_step": per_step}
    from oracle import restatement as R
    sd = synth.clip_state_dict(arch, 0)
    toks, n_ctx, _ = load_tokens()
    w = arch.transformer_width
    # text features: computed once, untimed — the reference caches them after the first call (T:421-439);
    # random unit rows stand in for them (their values do not change the timed image path's cost)
    g = torch.Generator().manual_seed(0)
    tfeat = [torch.nn.functional.normalize(torch.randn((toks.shape[0], arch.embed_dim), generator=g), dim=-1) for _ in range(3)]
    heads = arch.vision_width * 32 // 64

    def step(img):
        with torch.no_grad():
            feat = R.rn_trunk(sd, img, arch.vision_layers)
            return R.head_test(R.attnpool_global(sd, feat, heads), R.local_features(sd, feat), tfeat[0], tfeat[1], tfeat[2])

    probe = synth.images(1, arch.image_resolution, 99)
    step(probe)
    t0 = time.perf_counter()
    step(probe)
    t_img = time.perf_counter() - t0
    total_steps = steps + warmup
    per_step = max(1, min(32, int(budget_s / max(t_img, 1e-3) / max(total_steps, 1))))
    imgs = synth.images(per_step, arch.image_resolution, 100)
    for _ in range(warmup):
        step(imgs)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(imgs)
    dt = time.perf_counter() - t0
    return {"value": per_step * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} steps x {per_step} synthetic 448x448 images through oracle/restatement.py "
                      f"(RN101 trunk + attnpool + dual-prompt head, fp32, torch CPU, prompt features cached)",
            "ms_per_step": 1e3 * dt / steps, "images_per_step": per_step}


def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import synth
    arch = synth.RN101(448)
    r = cpu_reference_rate(arch, args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the product arm's config, verbatim (the bounded sample this arm actually ran is described in cpu_baseline.sample)
            "config": product_config(args.gpus, args.batch, 80),
            "bounded_sample": True, "sample_images_per_step": r["images_per_step"], "device": "host CPU",
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# the reference's own modules on the GPU (courtesy baseline): never part of the product path
# --------------------------------------------------------------------------------------------------
def gpu_reference_rate(arch, dev, batch, steps=5, warmup=2):
    """Eager PyTorch run of the reference's ModifiedResNet + the dense head of T:405-472 written with the reference's own
    tensor ops, from the staged copy of clip/model.py: the number a user of the reference sees on this box."""
    import importlib.util
    import torch.nn.functional as F
    path = os.path.join(ROOT, "baseline", "_ref", "project", "my_code", "clip", "model.py")
    if not os.path.exists(path):
        return {"unavailable": "baseline/_ref is not staged (run __graft_entry__.build() in the build container)"}
    from lecb200 import synth
    spec = importlib.util.spec_from_file_location("_ref_clip_model_gpu", path)
    M = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(M)
    clip = M.CLIP(*arch.ctor_args())
    clip.load_state_dict(synth.clip_state_dict(arch, 0), strict=False)
    visual = clip.visual.float().eval().to(dev)
    k, d = 80, arch.embed_dim
    g = torch.Generator().manual_seed(0)
    t_pos, t_neg, t_evi = (F.normalize(torch.randn((k, d), generator=g), dim=-1).to(dev) for _ in range(3))
    images = torch.randn((batch, 3, arch.image_resolution, arch.image_resolution), generator=g).to(dev)
    ap_ = visual.attnpool

    @torch.no_grad()
    def step():
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

    out = {"what": "reference PyTorch modules (staged clip/model.py + head of T:405-472), eager, same weights / batch / resolution",
           "batch": batch}
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    try:
        for name in ("bf16_autocast", "tf32"):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = name == "tf32"
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if name == "bf16_autocast" else torch.autocast("cuda", enabled=False)
            with ctx:
                for _ in range(warmup):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    del visual, clip, images
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------
# product arm
# --------------------------------------------------------------------------------------------------
class Ctx:
    """Rank / device / timing helpers shared by the headline and the secondary configurations."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup):
        """W untimed steps, barrier + sync, exactly K timed steps between two CUDA events, barrier + sync, max over ranks."""
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps, r


def build_clip(arch, dev):
    from lecb200 import synth
    from lecb200.clip_model import CLIPParams
    clip = CLIPParams(*arch.ctor_args())
    clip.load_state_dict(synth.clip_state_dict(arch, 0), strict=False)
    return clip.float().to(dev).eval()


def e2e_rate(cx, step_fn, host_in, dev_in, host_out, steps, warmup, pack):
    """Pinned host batch -> H2D on a copy stream (double-buffered) -> step -> D2H of the packed logits, all inside the timed
    region."""
    copy_stream = torch.cuda.Stream(device=cx.dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def loop(n):
        main_stream = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            dev_in[0].copy_(host_in[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(consumed[nxt])
                    dev_in[nxt].copy_(host_in[nxt], non_blocking=True)
                    ready[nxt].record(copy_stream)
            main_stream.wait_event(ready[cur])
            lg, ll = step_fn(dev_in[cur])
            consumed[cur].record(main_stream)
            host_out.copy_(pack(lg, ll), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    loop(warmup)
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop(steps)
    e1.record()
    cx.barrier()
    return cx.max_over_ranks(e0.elapsed_time(e1)) / steps


def bench_headline(cx, args):
    import lecb200
    from lecb200 import synth
    from lecb200.dense_clip import DenseCLIPB200
    from lecb200.dist import all_gather_logits, pack_logits
    from lecb200.prof import KernelTimer
    rank, world, dev = cx.rank, cx.world, cx.dev
    arch = synth.RN101(448)
    toks, n_ctx, names = load_tokens()
    model = DenseCLIPB200(make_cfg(448, n_ctx, True), names, build_clip(arch, dev), tokenized_prompts=toks).to(dev)
    B, K = args.batch, len(names)

    # ---------------- HBM-resident throughput ----------------
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    images = torch.randn((B, 3, 448, 448), device=dev, generator=gen)        # 616 MB > L2: no flush needed

    def step(img):
        out = model(img, if_test=True)
        return all_gather_logits(out[0], out[1])

    sampler = ClockSampler(cx.local_rank)
    if rank == 0:
        sampler.start()                      # started before warm-up: nvidia-smi needs ~0.5 s to produce samples
    for _ in range(args.warmup):
        step(images)
    cx.barrier()
    launches0 = lecb200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    if args.ncu_window:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        res = step(images)
    e1.record()
    cx.barrier()
    if args.ncu_window:
        torch.cuda.profiler.stop()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ms = cx.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = (lecb200.launch_count() - launches0) // args.steps
    value = world * B / (ms * 1e-3)
    assert torch.isfinite(res[0]).all() and torch.isfinite(res[1]).all()

    # ---------------- end to end: pinned host images -> logits on the host ----------------
    host_out = torch.empty((B * world, 2 * K), dtype=torch.float32).pin_memory()
    # (1) raw uint8 NHWC pixels (the decoder's output): 154 MB per 256 images
    host_u8 = [torch.randint(0, 256, (B, 448, 448, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    dev_u8 = [torch.empty((B, 448, 448, 3), device=dev, dtype=torch.uint8) for _ in range(2)]
    ms_u8 = e2e_rate(cx, step, host_u8, dev_u8, host_out, args.steps, args.warmup, pack_logits)
    del host_u8, dev_u8
    # (2) the reference's host tensor: normalised float NCHW, 616 MB per 256 images
    host_f = [torch.empty((B, 3, 448, 448), dtype=torch.float32).pin_memory() for _ in range(2)]
    for h in host_f:
        h.normal_()
    dev_f = [torch.empty((B, 3, 448, 448), device=dev) for _ in range(2)]
    ms_f = e2e_rate(cx, step, host_f, dev_f, host_out, max(3, args.steps // 2), args.warmup, pack_logits)
    del host_f, dev_f
    e2e = {"value": world * B / (ms_u8 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": world * B * 3 * 448 * 448,
           "d2h_bytes_per_step": world * B * 2 * K * 4, "ms_per_step": ms_u8,
           "input": "pinned host uint8 NHWC [B,448,448,3] raw pixels; ToTensor + Normalize inside lecb_stem_conv1_u8"}
    e2e_fp32 = {"value": world * B / (ms_f * 1e-3), "unit": UNIT, "h2d_bytes_per_step": world * B * 3 * 448 * 448 * 4,
                "d2h_bytes_per_step": world * B * 2 * K * 4, "ms_per_step": ms_f,
                "input": "pinned host fp32 NCHW [B,3,448,448], normalised (the reference DataLoader's tensor)"}

    # ---------------- roofline: per-entry-point CUDA events (instrumented replay, not the timed region) ----
    roofline, table = None, None
    if rank == 0:
        pk = peaks()
        prof_steps = 2
        with KernelTimer() as kt:
            for _ in range(prof_steps):
                model(images, if_test=True)
        table = kt.summary(prof_steps)
        fam = ("lecb_gemm_bf16", "lecb_gemm_bf16_dual", "lecb_conv3x3_bf16")
        gemm_ms = sum(table[n]["ms"] for n in fam if n in table)
        gemm_fl = sum(table[n]["flops"] for n in fam if n in table)
        gemm_n = sum(table[n]["launches"] for n in fam if n in table)
        total_ms = sum(d["ms"] for d in table.values())
        achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12
        # DRAM bytes per launch of the same kernel family, from the committed ncu pass over one bench step
        # (tools/gpu_profile_traffic.sh -> tools/summarize_ncu.py); never measured live (ncu is not a bench)
        traffic, traffic_src = None, None
        for tname in ("r02_step_traffic_summary.json", "r01_step_traffic_summary.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath) and B == 256:
                with open(tpath) as f:
                    fams = json.load(f)["by_kernel_family"]
                parts = [fams[k] for k in ("gemm_kernel", "gemm_pair_kernel") if k in fams]      # single-CTA and CTA-pair kernels
                if parts:
                    n_l = sum(x["launches"] for x in parts)
                    traffic = sum(x["dram_bytes_total"] for x in parts) / max(n_l, 1)
                    traffic_src = f"profiles/{tname} (ncu dram__bytes_read+write, mean over the step's {n_l} GEMM / conv launches)"
                    break
        gemm_by = sum(table[n]["bytes"] for n in fam if n in table)
        # The kernel serves layers on both sides of the ridge: per (entry point, shape) the bound is
        # max(flops / tensor peak, algorithmic bytes / HBM peak); their sum over the step vs the measured time says
        # how close the family is to its own per-layer rooflines ("report per-layer max(...)").
        t_bound = t_meas = t_hbm_bound_layers = 0.0
        for r in kt.detail(prof_steps):
            if r["op"] not in fam:
                continue
            ms_r = r["ms_per_step"]
            tb_t = r["tflops"] * ms_r / pk["bf16_sustained"]
            tb_h = r["gbs"] * ms_r / pk["hbm_gbs"]
            t_bound += max(tb_t, tb_h)
            t_meas += ms_r
            if tb_h > tb_t:
                t_hbm_bound_layers += ms_r
        roofline = {"kernel": "lecb::gemm_kernel<BN,BK,conv> (tcgen05 GEMM + TMA-im2col conv)", "bound": "tensor",
                    "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                    "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": gemm_by / max(gemm_n, 1),
                    "per_layer_bound": {"frac": t_bound / max(t_meas, 1e-9), "ms_at_bound": t_bound, "ms_measured": t_meas,
                                        "time_share_of_hbm_bound_layers": t_hbm_bound_layers / max(t_meas, 1e-9),
                                        "hbm_peak_gbs": pk["hbm_gbs"],
                                        "note": "sum over (entry point, shape) of max(flops/tensor peak, algorithmic bytes/HBM peak) / measured"},
                    "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)",
                    "launches_per_step": gemm_n, "avg_launch_us": 1e3 * gemm_ms / max(gemm_n, 1),
                    "share_of_step": gemm_ms / total_ms, "algorithmic_gflop_per_step": gemm_fl / 1e9}
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"batch": B, "ms_per_step_events_sum": total_ms, "table": table, "detail": kt.detail(prof_steps)}, f, indent=1)
    del model, images
    torch.cuda.empty_cache()
    return dict(value=value, ms=ms, launches=int(launches), clocks=clocks, e2e=e2e, e2e_fp32=e2e_fp32, roofline=roofline,
                table=table, K=K)


def bench_vit(cx, name, batch, steps, warmup):
    """BASELINE configs[2] / [4]: ViT-B/16 / ViT-L/14 dual-prompt inference at 448x448, batch-sharded, logits all-gathered."""
    import lecb200
    from lecb200 import synth
    from lecb200.dense_clip import DenseCLIPB200
    from lecb200.dist import all_gather_logits
    arch = synth.VITB16(448) if name == "vitb16" else synth.VITL14(448)
    toks, n_ctx, names = load_tokens()
    model = DenseCLIPB200(make_cfg(448, n_ctx, True), names, build_clip(arch, cx.dev), tokenized_prompts=toks).to(cx.dev)
    images = torch.randn((batch, 3, 448, 448), device=cx.dev)

    def step():
        out = model(images, if_test=True)
        return all_gather_logits(out[0], out[1])

    n0 = lecb200.launch_count()
    ms, res = cx.timed(step, steps, warmup)
    launches = (lecb200.launch_count() - n0) // (steps + warmup)
    gf_img = 156.99 + 1.5 if name == "vitb16" else 723.59 + 3.0          # SURVEY 8d, algorithmic GF per image
    pk = peaks()
    out = {"metric": METRIC, "value": cx.world * batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
           "config": {"workload": f"CLIP {'ViT-B/16' if name == 'vitb16' else 'ViT-L/14'} dual-prompt inference, 80 classes, 448x448, "
                                  f"batch {batch} per GPU, logits all-gathered", "global_batch": cx.world * batch,
                      "parallelism": f"dp{cx.world}"},
           "tflops_per_gpu_algorithmic": gf_img * batch / ms, "frac_of_sustained_bf16_peak": gf_img * batch / ms / pk["bf16_sustained"],
           "gpu_launches": int(launches), "finite": bool(torch.isfinite(res[0]).all() and torch.isfinite(res[1]).all()),
           "parity": "global feature pinned to the reference VisionTransformer; dense head vs repo oracle (the reference has no ViT dense path)"}
    del model, images
    torch.cuda.empty_cache()
    return out


def bench_train(cx, per_gpu_batch, steps, warmup, use_graph=True):
    """BASELINE configs[3]: text-only prompt-tuning step (CLIP RN50 text tower + ASL), 77-token synthetic captions with the
    real caption-length distribution, prompt-gradient all-reduce, SGD step."""
    import lecb200
    from lecb200 import losses, synth
    from lecb200.dense_clip import DenseCLIPB200
    from lecb200.dist import PromptSGD, broadcast_params
    dev, rank, world = cx.dev, cx.rank, cx.world
    arch = synth.RN50(224)
    toks, n_ctx, names = load_tokens()
    model = DenseCLIPB200(make_cfg(224, n_ctx, False), names, build_clip(arch, dev), tokenized_prompts=toks).to(dev)
    for n_, p in model.named_parameters():
        p.requires_grad_("prompt_learner." in n_ and "prompt_learner_m" not in n_)
    params = [p for p in model.prompt_learner.parameters()]
    broadcast_params(params)                 # DDP's construction-time broadcast (T:786-787)
    opt = PromptSGD(params, lr=0.002, momentum=0.9)
    b = per_gpu_batch
    caps = synth.captions(b, 100 + rank, vocab=arch.vocab_size)
    # the captured step cannot read the batch's longest caption back from the device: promise it up front
    model.caption_len_hint = int(caps.argmax(-1).max()) + 1
    caps = caps.to(dev)
    y = synth.labels(b, len(names), 100 + rank).to(dev)

    def compute():                           # forward + loss + backward + gradients packed into the flat bucket
        out = model(None, caps)
        loss = losses.ASL_loss(out[0], y) + losses.ASL_loss(out[1], y)
        opt.zero_grad()
        loss.backward()
        opt.pack()
        return loss

    def eager_step():
        loss = compute()
        opt.reduce_and_update()              # all-reduce (N > 1) -> fused SGD from the flat bucket
        return loss

    step, graphed = eager_step, False
    for _ in range(3):
        eager_step()
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    eager_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad()
            # one rank: the whole step is one graph.  Several ranks: the collective and the update stay outside the capture
            # (two launches + one all-reduce per step), so a capture never wraps an NCCL call
            with torch.cuda.graph(graph):
                static_loss = eager_step() if world == 1 else compute()

            def step():
                graph.replay()
                if world > 1:
                    opt.reduce_and_update()
                return static_loss
            step()
            graphed = True
        except Exception as e:  # pragma: no cover
            print(f"[bench] CUDA-graph capture of the prompt-tuning step failed ({type(e).__name__}: {e}); timing eager",
                  file=sys.stderr)
            step = eager_step
            torch.cuda.synchronize()
    n0 = lecb200.launch_count()
    ms, loss = cx.timed(step, steps, warmup)
    launches = (lecb200.launch_count() - n0) // (steps + warmup)
    n0 = lecb200.launch_count()
    eager_step()
    launches_eager = lecb200.launch_count() - n0
    out = {"metric": "prompt_tuning_captions_per_sec", "value": world * b / (ms * 1e-3), "unit": "captions/s", "ms_per_step": ms,
           "config": {"workload": "text-only prompt-tuning step: CLIP RN50 text tower, 77-token synthetic captions (real length "
                                  "distribution), 160 prompt sequences, ASL loss, prompt-gradient all-reduce, SGD",
                      "per_gpu_batch": b, "global_batch": world * b, "parallelism": f"dp{world}"},
           "cuda_graph": graphed, "gpu_launches": int(launches if not graphed else launches_eager),
           "caption_positions_run": int(model.caption_len_hint), "caption_positions_total": int(caps.shape[1]),
           "final_loss": float(loss.detach()), "finite": bool(torch.isfinite(loss.detach()))}
    del model
    torch.cuda.empty_cache()
    return out


def sharded_prompt_check(cx):
    """Untimed self-check at N > 1 (driver-witnessed): the class-sharded prompt branch must reproduce the replicated one —
    logits of this rank's captions and the prompt gradients after the flat average."""
    from lecb200 import losses, synth
    from lecb200.dense_clip import DenseCLIPB200
    from lecb200.dist import allreduce_mean_grads, broadcast_params
    dev, rank = cx.dev, cx.rank
    arch = synth.RN50(224)
    toks, n_ctx, names = load_tokens()
    model = DenseCLIPB200(make_cfg(224, n_ctx, True), names, build_clip(arch, dev), tokenized_prompts=toks).to(dev)
    for n_, p in model.named_parameters():
        p.requires_grad_("prompt_learner." in n_ and "prompt_learner_m" not in n_)
    params = [p for p in model.prompt_learner.parameters()]
    broadcast_params(params)
    caps = synth.captions(16, 300 + rank, vocab=arch.vocab_size).to(dev)
    y = synth.labels(16, len(names), 300 + rank).to(dev)
    res = {}
    for mode in (False, True, None):          # None: the replicated branch a second time (its own run-to-run spread)
        model.shard_prompt_branch = bool(mode)
        for p in params:
            p.grad = None
        out = model(None, caps)
        loss = losses.ranking_loss(out[0], y, scale_=1.0, margin_=1) + losses.ranking_loss(out[1], y, scale_=1.0, margin_=1)
        loss.backward()
        allreduce_mean_grads(params)
        res[mode] = (out[0].detach().clone(), out[1].detach().clone(),
                     torch.cat([(torch.zeros_like(p) if p.grad is None else p.grad).reshape(-1) for p in params]))
    d_logits = max(float((res[True][0] - res[False][0]).abs().max()), float((res[True][1] - res[False][1]).abs().max()))
    g0, g1, g0b = res[False][2], res[True][2], res[None][2]
    gmax, gnorm = g0.abs().max().clamp_min(1e-12), g0.norm().clamp_min(1e-12)
    d_grad = float((g1 - g0).abs().max() / gmax)
    d_l2 = float((g1 - g0).norm() / gnorm)
    cos = float(torch.nn.functional.cosine_similarity(g0, g1, dim=0))
    self_max = float((g0b - g0).abs().max() / gmax)          # replicated vs replicated: fp32 atomics feed bf16 roundings
    self_l2 = float((g0b - g0).norm() / gnorm)
    t = torch.tensor([d_logits, d_grad, 1.0 - cos, d_l2, self_max, self_l2], device=dev, dtype=torch.float64)
    cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
    del model
    torch.cuda.empty_cache()
    d_logits, d_grad, d_cos, d_l2, self_max, self_l2 = (float(v) for v in t)
    return {"what": "class-sharded prompt branch vs replicated branch, same captions (max over ranks)",
            "logits_max_abs_diff": d_logits, "grad_rel_l2": d_l2, "grad_one_minus_cosine": d_cos,
            "grad_max_diff_over_max": d_grad,
            "replicated_vs_itself": {"grad_rel_l2": self_l2, "grad_max_diff_over_max": self_max},
            # The forward is deterministic (logits gate 1e-3 absolute, measured 1e-6).  The backward is not bit-reproducible
            # (fp32 atomics feeding bf16 roundings: `replicated_vs_itself` is the same statistic between two runs of ONE
            # branch), and the two branches round at different points (sum of the ranks' feature gradients, then twelve bf16
            # layers, against twelve bf16 layers per rank, then the average).  Gated: the vector statistics (relative L2 error
            # 5 %, cosine > 0.999).  The largest single-element deviation is reported, not gated: over four driver / builder
            # runs it was 0.031-0.054 of the largest gradient element, i.e. it straddles the 5 % the first version gated on
            "tolerance": {"logits": 1e-3, "grad_rel_l2": 5e-2, "grad_one_minus_cosine": 1e-3},
            "ok": bool(d_logits <= 1e-3 and d_l2 <= 5e-2 and d_cos <= 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lecb200", choices=["lecb200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step (configs[1]: 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configurations (ViT towers, prompt tuning)")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-entry-point timing table (JSON) here")
    ap.add_argument("--ncu-window", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (use with ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    args.warmup = max(args.warmup, 3)
    cx = Ctx()
    rank, world = cx.rank, cx.world
    h = bench_headline(cx, args)

    extra = {}
    if not args.no_extra:
        k2 = max(5, args.steps // 2)
        for name, fn in (("vitb16", lambda: bench_vit(cx, "vitb16", 128, k2, 3)),
                         ("vitl14", lambda: bench_vit(cx, "vitl14", 128, max(3, args.steps // 4), 3)),
                         ("prompt_tuning", lambda: bench_train(cx, 64, max(20, args.steps), 5))):
            try:
                extra[name] = fn()
            except Exception as e:  # a secondary line must never take the headline down
                extra[name] = {"error": f"{type(e).__name__}: {e}"}
                torch.cuda.empty_cache()
        if world > 1:
            try:
                extra["sharded_prompt_branch_check"] = sharded_prompt_check(cx)
            except Exception as e:
                extra["sharded_prompt_branch_check"] = {"error": f"{type(e).__name__}: {e}"}

    gpu_reference = None
    if rank == 0 and world == 1 and not args.no_gpu_reference:
        from lecb200 import synth
        try:
            gpu_reference = gpu_reference_rate(synth.RN101(448), cx.dev, args.batch)
        except Exception as e:
            gpu_reference = {"error": f"{type(e).__name__}: {e}"}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from lecb200 import synth
        r = cpu_reference_rate(synth.RN101(448), steps=3, warmup=1, budget_s=25.0)
        cpu_baseline = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": h["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": h["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": product_config(world, args.batch, h["K"]),
                "clocks": h["clocks"], "e2e": h["e2e"], "e2e_fp32": h["e2e_fp32"], "gpu_launches": h["launches"],
                "roofline": h["roofline"], "cpu_baseline": cpu_baseline, "gpu_reference": gpu_reference, "extra": extra}
        if h["table"] is not None:
            line["kernel_ms_per_step"] = {k: round(v["ms"], 4) for k, v in sorted(h["table"].items(), key=lambda kv: -kv[1]["ms"])}
        print(json.dumps(line), flush=True)
    if world > 1:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
