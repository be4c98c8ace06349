#!/bin/bash
# round-2 call 1: run the artefacts round 1 left unrun + the courtesy GPU baseline of the reference modules
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
echo "== mma_pair" ; timeout 60 ./tools/micro/mma_pair > gpurun_out/c1_mma_pair.txt 2>&1; echo "rc=$?" >> gpurun_out/c1_mma_pair.txt; cat gpurun_out/c1_mma_pair.txt
echo "== crop_resize" ; timeout 120 python tools/micro/crop_resize_check.py > gpurun_out/c1_crop_resize.txt 2>&1; echo "rc=$?" >> gpurun_out/c1_crop_resize.txt; tail -5 gpurun_out/c1_crop_resize.txt
echo "== reference on the GPU"
for dt in bf16 tf32 fp32; do
  timeout 300 python tools/bench_reference_gpu.py --batch 64 --dtype $dt --steps 5 --warmup 2 >> gpurun_out/c1_reference_gpu.jsonl 2>> gpurun_out/c1_reference_gpu.err
done
timeout 300 python tools/bench_reference_gpu.py --batch 256 --dtype bf16 --steps 5 --warmup 2 >> gpurun_out/c1_reference_gpu.jsonl 2>> gpurun_out/c1_reference_gpu.err
cat gpurun_out/c1_reference_gpu.jsonl; tail -3 gpurun_out/c1_reference_gpu.err
