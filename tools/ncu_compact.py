"""Compact table of the ncu raw pages exported on the GPU box (tools/gpu_calls/gpu_call8.sh): python tools/ncu_compact.py gpurun_out/c8_*_raw.csv"""
import csv
import sys

WANT = [("time_us", "gpu__time_duration.sum"), ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"), ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tmem_active_pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("tensor_pipe_realtime_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
        ("grid", "launch__grid_size"), ("lts_sectors", "lts__t_sectors.sum"), ("sm_hz", "sm__cycles_elapsed.avg.per_second")]
for f in sys.argv[1:]:
    r = list(csv.reader(open(f)))
    if len(r) < 3:
        continue
    d = dict(zip(r[0], zip(r[2], r[1])))
    name = d.get("Kernel Name", ("?", ""))[0][:60]
    row = {"file": f.split("/")[-1].replace("_raw.csv", ""), "kernel": name}
    for k, m in WANT:
        if m in d:
            try:
                row[k] = round(float(d[m][0].replace(",", "")), 2)
            except ValueError:
                row[k] = d[m][0]
            if k in ("time_us", "dram_rd_MB", "dram_wr_MB", "sm_hz"):
                row[k + "_unit"] = d[m][1]
    print(row)
