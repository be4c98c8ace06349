"""How far does the REFERENCE adapter model's own prompt gradient move under rounding-level perturbations?  (CPU, fp32.)
The adapter's two ReLUs (Caption_distill_double_adapter.py:304-317) make d loss / d ctx discontinuous in the adapter's input, so
the tolerance of tests/test_round2_gpu.py::test_adapter_model_matches_reference is set from these numbers, measured on the golden
case (tests/golden/adapter_rn50.npz) with the fp32 restatement (pinned to the reference classes by tests/test_oracle_ext.py):

    perturbation of the adapter input / operands        max |dg| / max |g|      cosine
    gaussian, 1e-3 of mean |x|                          0.05 - 0.11             0.9953 - 0.9991
    gaussian, 5e-3 of mean |x|                          0.14 - 0.18             0.987  - 0.990
    adapter operands rounded to bf16                    0.09 - 0.11             0.9955 - 0.9963
    logits perturbed by 5e-3 (hinges only)              0                       1
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import restatement as R  # noqa: E402
from tests.test_oracle_ext import adapter_case  # noqa: E402

c = adapter_case()
arch, sd = c["arch"], c["sd"]
wd, wu = c["adapter"]


def run(noise, mode):
    pl = {k: (v.clone().requires_grad_(True) if k in ("ctx", "ctx_double") else v) for k, v in c["pl_state"].items()}
    seq = R.text_encode(sd, R.embed_tokens(sd, c["captions"]), None, arch.transformer_heads, sequence=True)
    eot = c["tokens"].argmax(dim=-1)
    feats = []
    for k in ("ctx", "ctx_double"):
        x = R.assemble_prompts(pl["token_prefix"], pl[k], pl["token_suffix"]) + sd["positional_embedding"]
        mask = torch.full((x.shape[1], x.shape[1]), float("-inf")).triu_(1)
        i = 0
        while f"transformer.resblocks.{i}.ln_1.weight" in sd:
            x = R._res_block(sd, f"transformer.resblocks.{i}", x, arch.transformer_heads, mask)
            i += 1
        torch.manual_seed(3)
        if mode == "input":
            x = x + noise * x.detach().abs().mean() * torch.randn_like(x)
        if mode == "bf16":
            exact = torch.relu(x @ wd.t())
            rounded = torch.relu(x.bfloat16().float() @ wd.bfloat16().float().t()).bfloat16().float()
            x = x + torch.relu((exact + (rounded - exact).detach()) @ wu.bfloat16().float().t())
        else:
            x = x + torch.relu(torch.relu(x @ wd.t()) @ wu.t())
        x = F.layer_norm(x, (x.shape[-1],), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
        feats.append(x[torch.arange(x.shape[0]), eot] @ sd["text_projection"])
    r = R.head_train(seq, c["captions"], feats[0], feats[1], None, 4.0, 50.0)
    loss = R.ranking_loss(r[0], c["labels"], 1.0, 1.0) + R.ranking_loss(r[1], c["labels"], 1.0, 1.0)
    loss.backward()
    return pl["ctx"].grad.numpy(), pl["ctx_double"].grad.numpy()


base = run(0.0, "none")
for noise, mode in ((1e-3, "input"), (5e-3, "input"), (0.0, "bf16")):
    for name, a, b in zip(("ctx", "ctx_double"), run(noise, mode), base):
        cos = float(a.flatten() @ b.flatten() / np.linalg.norm(a) / np.linalg.norm(b))
        print(f"{mode:6s} {noise:7.0e} {name:10s} max err / max = {np.abs(a - b).max() / np.abs(b).max():.4f}  cosine = {cos:.5f}")
