#!/bin/bash
# Per-launch DRAM traffic + duration of one steady-state bench step (metrics-only ncu pass), and the same for
# the row-kernel roofline bench.  Outputs: gpurun_out/<tag>_step_traffic.csv, <tag>_rowops_traffic.csv
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ncu-window"
$CMD > gpurun_out/${TAG}_traffic_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_traffic_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off -c 200 --csv --log-file gpurun_out/${TAG}_step_traffic.csv $CMD > gpurun_out/${TAG}_traffic_ncu.log 2>&1
echo "step traffic rc=$?"
python tools/bench_rowops.py > gpurun_out/${TAG}_rowops.jsonl 2>&1 || { echo "rowops failed"; tail -5 gpurun_out/${TAG}_rowops.jsonl; }
cat gpurun_out/${TAG}_rowops.jsonl
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'head_aggregate|asl_fwd_bwd|l2norm_kernel|ranking|layernorm_fwd|avgpool' -c 60 --csv \
    --log-file gpurun_out/${TAG}_rowops_traffic.csv python tools/bench_rowops.py > gpurun_out/${TAG}_rowops_ncu.log 2>&1
echo "rowops traffic rc=$?"
