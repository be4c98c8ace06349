"""One conv3x3 shape, a few launches (target for an ncu --set full capture): python conv_one.py B H W Cin Cout"""
import sys
import torch
from lecb200 import ops

b, h, w, ci, co = [int(v) for v in sys.argv[1:6]]
x = torch.randn((b, h, w, ci), device="cuda").bfloat16()
wt = (torch.randn((co, 3, 3, ci), device="cuda") * (9 * ci) ** -0.5).bfloat16()
bias = torch.randn((co,), device="cuda")
out = torch.empty((b, h, w, co), device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.conv3x3(x, wt, bias, out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    ops.conv3x3(x, wt, bias, out=out)
e.record()
torch.cuda.synchronize()
print(f"conv {b}x{h}x{w} {ci}->{co}: {s.elapsed_time(e) / 5:.3f} ms")
