"""Summarise where the warp roles of a warp-specialised kernel wait: python tools/ncu_waits.py report.ncu-rep [launch index]
Prints every mbarrier try-wait / async-op SASS instruction with its stall samples (next instruction included)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = heads[which]
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
hdr = rows[h]
data = [r for r in rows[h + 1:end] if len(r) == len(hdr)]
rows = rows[h - 1:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data)
print(rows[0][1][:120], "total samples", tot)
keys = ("SYNCS.PHASECHK", "UTCHMMA", "UTMALDG", "UTCBAR", "UTMASTG", "BAR.SYNC", "LDTM", "UTMACMDFLUSH", "DEPBAR")
mma = 0
for i, r in enumerate(data):
    s = r[isrc]
    if any(k in s for k in keys):
        n = int(r[isamp] or 0) + (int(data[i + 1][isamp] or 0) if i + 1 < len(data) else 0)
        if n or int(r[iex] or 0):
            print(f"{i:5d} samples {n:6d} ({100 * n / tot:4.1f}%) exec {r[iex]:>9s}  {s[:96]}")
