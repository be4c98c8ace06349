"""Build a caption feature bank at the reference's size (220k synthetic captions with the shipped length distribution,
RN50 text tower -> [220000,1024] fp16) and report captions/s.  `--out PATH` writes the reference's pickle format."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_tokens, make_cfg  # noqa: E402
from lecb200 import bank as B  # noqa: E402
from lecb200 import synth  # noqa: E402
from lecb200.clip_model import CLIPParams  # noqa: E402
from lecb200.dense_clip import TextEncoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=220000)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--out", default=None)
args = ap.parse_args()
arch = synth.RN50(224)
clip = CLIPParams(*arch.ctor_args())
clip.load_state_dict(synth.clip_state_dict(arch, 0), strict=False)
enc = TextEncoder(clip.float().cuda().eval())
base = synth.captions(8192, 7, vocab=arch.vocab_size)
caps = base.repeat((args.n + 8191) // 8192, 1)[:args.n]
B.build_caption_bank(enc, caps[:args.batch], args.batch)          # warm-up (engine construction, kernels)
torch.cuda.synchronize()
t0 = time.perf_counter()
bank = B.build_caption_bank(enc, caps, args.batch)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
norms = bank.float().norm(dim=-1)
print(json.dumps({"op": "caption bank build", "captions": args.n, "shape": list(bank.shape), "dtype": str(bank.dtype),
                  "seconds": round(dt, 3), "captions_per_sec": round(args.n / dt, 1),
                  "unit_norm_max_dev": float((norms - 1).abs().max())}))
if args.out:
    B.save_caption_bank(args.out, bank)
    print("wrote", args.out, os.path.getsize(args.out), "bytes")
