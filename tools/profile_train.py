"""Per-entry-point CUDA-event breakdown of one prompt-tuning step (see tools/bench_train.py)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_tokens, make_cfg  # noqa: E402
from lecb200 import losses, synth  # noqa: E402
from lecb200.clip_model import CLIPParams  # noqa: E402
from lecb200.dense_clip import DenseCLIPB200  # noqa: E402
from lecb200.prof import KernelTimer  # noqa: E402

dev = torch.device("cuda", 0)
arch = synth.RN50(224)
toks, n_ctx, names = load_tokens()
clip = CLIPParams(*arch.ctor_args())
clip.load_state_dict(synth.clip_state_dict(arch, 0), strict=False)
clip = clip.float().to(dev).eval()
model = DenseCLIPB200(make_cfg(224, n_ctx, False), names, clip, tokenized_prompts=toks).to(dev)
caps = synth.captions(64, 100, vocab=arch.vocab_size).to(dev)
y = synth.labels(64, len(names), 100).to(dev)


def step():
    out = model(None, caps)
    loss = losses.ASL_loss(out[0], y) + losses.ASL_loss(out[1], y)
    for p in model.prompt_learner.parameters():
        p.grad = None
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with KernelTimer() as kt:
    for _ in range(3):
        step()
rows = kt.detail(3)
tot = sum(r["ms_per_step"] for r in rows)
print("sum of lecb kernels per step: %.3f ms" % tot)
agg = {}
for r in rows:
    agg[r["op"]] = agg.get(r["op"], 0) + r["ms_per_step"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{k:28s} {v:8.3f} ms")
for r in rows[:12]:
    print(r)
