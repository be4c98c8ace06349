#!/bin/bash
# Round-2 ncu evidence for the bench step (run on the GPU box via gpurun): launch list (device time per launch) and DRAM
# traffic per launch of one steady-state step of the SAME command; summarised on the box (reports are not brought back).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-gpu-reference --ncu-window"
$CMD > gpurun_out/r02_traffic_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_traffic_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off -c 140 --csv --log-file gpurun_out/r02_step_traffic.csv $CMD > gpurun_out/r02_traffic_ncu.log 2>&1
echo "step traffic rc=$?"
python tools/summarize_ncu.py gpurun_out/r02_step_traffic.csv gpurun_out/r02_step_traffic_summary.json
