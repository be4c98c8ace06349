#!/bin/bash
# ncu --set full on nine consecutive layer3 launches (3 bottlenecks: 1x1 reduce, 3x3, 1x1 expand+residual).
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --ncu-window"
mkdir -p gpurun_out
$CMD > gpurun_out/l3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -s 50 -c 9 \
    -o gpurun_out/r01_layer3_full $CMD > gpurun_out/l3_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/l3_ncu.log
