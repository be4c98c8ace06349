"""Caption retrieval (T:444-448) at the reference's bank size: 220 000 x 1024 fp16 (RN50 embed dim),
256 queries.  Prints per-kernel CUDA-event times and the effective bandwidth over the bank."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lecb200 import synth  # noqa: E402
from lecb200.prof import KernelTimer  # noqa: E402
from lecb200.retrieval import retrieve_mean  # noqa: E402

n, d, b = 220000, 1024, 256
bank = synth.caption_bank(n, d, 0).cuda()
q = torch.nn.functional.normalize(torch.randn((b, d), device="cuda"), dim=-1)
for _ in range(3):
    retrieve_mean(q, bank)
torch.cuda.synchronize()
blocks = []                       # median of five 10-call blocks: a single block is at the mercy of the host (allocator, page-ins)
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g_add, vals = retrieve_mean(q, bank)
    e1.record()
    torch.cuda.synchronize()
    blocks.append(e0.elapsed_time(e1) / 10)
ms = sorted(blocks)[2]
with KernelTimer() as kt:
    retrieve_mean(q, bank)
rows = {r["op"] + (" " + r["shape"] if r["shape"] else ""): round(r["ms_per_step"], 4) for r in kt.detail(1)}
sim = q @ bank.float().t()
ref_v, ref_i = sim.topk(10, -1)
print(json.dumps({"op": "caption retrieval", "bank": [n, d], "queries": b, "ms": round(ms, 3), "ms_blocks": [round(x, 3) for x in blocks], "kernels_ms": rows,
                  "bank_GBs": round(n * d * 2 * 2 / ms / 1e6, 1), "sim_matrix_MB": b * n * 4 / 1e6,
                  "topk_score_max_err": float((vals - ref_v).abs().max())}))
