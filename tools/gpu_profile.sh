#!/bin/bash
# Run on the GPU box (via gpurun): plain bench run, ncu launch list of the same command, one --set full
# capture of the dominant kernel.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
# launch list: skip prompt encoding + warm-up (3*87 text-tower launches + 3 steps), record two steady steps
ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 250 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 600 -c 12 \
    -o gpurun_out/${TAG}_gemm_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
