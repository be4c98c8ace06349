"""Condense an `ncu --page raw --csv` export of a --set full capture into the few columns DESIGN.md cites.

    ncu -i gpurun_out/r01f_gemm_full.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_full_summary.py raw.csv profiles/r01_gemm_kernel_ncu_full_summary.csv "header comment"
"""
import csv
import sys

WANT = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg"]
src, dst = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = [hdr.index(w) for w in WANT if w in hdr]
with open(dst, "w", newline="") as f:
    f.write(f'"# {note}; units: ' + ", ".join(f"{hdr[i].split('.')[0]}={units[i]}" for i in idx if units[i]) + '"\n')
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx])
    for r in data:
        w.writerow([r[i] for i in idx])
print(f"{len(data)} launches -> {dst}")
