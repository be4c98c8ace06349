"""Summarise an ncu metrics CSV (gpu__time_duration / dram__bytes_read / dram__bytes_write per launch) per kernel.

    python tools/summarize_ncu.py gpurun_out/r01_step_traffic.csv profiles/r01_step_traffic_summary.json
"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict


PCT = "dram__throughput.avg.pct_of_peak_sustained_elapsed"


def short(name):
    """'void lecb::gemm_kernel<128, 64, ...>(args)' or ncu's namespace-less 'void l2norm_kernel<float, float, 4>(args)' -> the
    kernel's own name; template arguments are kept (GEMM instances and row-kernel variants are different kernels)."""
    head = name.split("(")[0]
    m = re.search(r"(\w+)\s*(<.*>)?\s*$", head)
    if not m:
        return head[-60:]
    return m.group(1) + (m.group(2) or "")


def main(src, dst):
    launches = OrderedDict()
    with open(src) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[1:]:
        k = launches.setdefault(r[idx["ID"]], {"kernel": r[idx["Kernel Name"]], "grid": r[idx["Grid Size"]]})
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        name = r[idx["Metric Name"]]
        if name == "gpu__time_duration.sum":
            v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)      # -> us
        elif unit == "Kbyte":
            v *= 1e3
        elif unit == "Mbyte":
            v *= 1e6
        elif unit == "Gbyte":
            v *= 1e9
        k[name] = v
    agg = defaultdict(lambda: {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0, "pct": 0.0, "pct_n": 0})
    fam = defaultdict(lambda: {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0, "pct": 0.0, "pct_n": 0})
    for l in launches.values():
        for table, key in ((agg, short(l["kernel"])), (fam, re.sub(r"<.*", "", short(l["kernel"])))):
            a = table[key]
            a["launches"] += 1
            a["us"] += l.get("gpu__time_duration.sum", 0.0)
            a["dram_read"] += l.get("dram__bytes_read.sum", 0.0)
            a["dram_write"] += l.get("dram__bytes_write.sum", 0.0)
            if PCT in l:                                  # optional fourth metric of the row-kernel pass
                a["pct"] += l[PCT]
                a["pct_n"] += 1
    total_us = sum(a["us"] for a in fam.values())

    def fin(t):
        out = {}
        for k, a in sorted(t.items(), key=lambda kv: -kv[1]["us"]):
            by = a["dram_read"] + a["dram_write"]
            out[k] = {"launches": a["launches"], "us_total": round(a["us"], 1), "share_of_profiled_time": round(a["us"] / total_us, 4),
                      "dram_bytes_total": by, "dram_bytes_per_launch": by / a["launches"],
                      "dram_GBs_under_ncu": round(by / a["us"] / 1e3, 1) if a["us"] else None}
            if a["pct_n"]:
                out[k]["dram_throughput_pct_of_peak_mean"] = round(a["pct"] / a["pct_n"], 1)
        return out

    res = {"source": src, "note": "ncu replays each kernel cold-cache and serialised: shares and bytes are meaningful, absolute times are not bench values",
           "launches_profiled": len(launches), "by_kernel_family": fin(fam), "by_kernel_instance": fin(agg)}
    with open(dst, "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res["by_kernel_family"].items():
        print(f"{k:32s} n={v['launches']:4d} us={v['us_total']:10.1f} share={v['share_of_profiled_time']:.3f} dram/launch={v['dram_bytes_per_launch']/1e6:9.1f} MB")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
