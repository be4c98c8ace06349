"""Per-source-line share of executed warp instructions and stall samples of one kernel in an ncu report compiled with -lineinfo:
python tools/ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file, hdr = None, None
per_line, samples, src = collections.Counter(), collections.Counter(), {}


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8 or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (cur_file, ln)
    per_line[key] += num(r[7])
    samples[key] += num(r[6])
    src[key] = r[1][:110]
tot, ts = sum(per_line.values()) or 1, sum(samples.values()) or 1
print(f"total warp instructions {tot}, stall samples {ts}")
for k, v in per_line.most_common(top):
    print(f"{v / tot * 100:5.1f}% instr {samples[k] / ts * 100:5.1f}% samples  {k[0]}:{k[1]}  {src[k]}")
