#!/bin/bash
# Run on the GPU box (via gpurun): plain bench run, ncu launch list of the same command, one --set full
# capture of the dominant kernel.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ncu-window"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
# launch list of the timed region only (cudaProfilerStart/Stop window): two steady-state steps
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 300 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -s 30 -c 12 \
    -o gpurun_out/${TAG}_gemm_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
