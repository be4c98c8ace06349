"""Compare per-shape step profiles (bench.py --profile-out): python tools/cmp_prof.py a.json b.json ..."""
import json
import sys

files = sys.argv[1:]
tabs = []
for f in files:
    d = json.load(open(f))
    tabs.append(({(r["op"], r["shape"]): r["ms_per_step"] for r in d["detail"]}, d["ms_per_step_events_sum"]))
keys = sorted(set().union(*[t[0].keys() for t in tabs]), key=lambda k: -max(t[0].get(k, 0) for t in tabs))
print(" " * 62 + " ".join(f"{f.split('/')[-1][:10]:>10s}" for f in files))
print(f"{'TOTAL':62s}" + " ".join(f"{t[1]:10.3f}" for t in tabs))
for k in keys[:int(__import__('os').environ.get('TOP', 24))]:
    print(f"{(k[0][5:] + ' ' + k[1])[:62]:62s}" + " ".join(f"{t[0].get(k, float('nan')):10.3f}" for t in tabs))
