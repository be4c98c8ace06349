#!/bin/bash
# N-GPU refresh (gpurun --gpus N): headline bench, ViT benches and the prompt-tuning step under torchrun
set -u
N=$1; TAG=${2:-r01f}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; tail -1 gpurun_out/${TAG}_bench_${N}gpu.json | cut -c1-200
timeout 600 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference_${N}gpu.json 2>> gpurun_out/${TAG}_bench_${N}gpu.err; tail -1 gpurun_out/${TAG}_bench_reference_${N}gpu.json | cut -c1-200
timeout 400 $TR tools/bench_vit.py --arch vitb16 > gpurun_out/${TAG}_vitb16_${N}gpu.json 2>> gpurun_out/${TAG}_bench_${N}gpu.err; tail -1 gpurun_out/${TAG}_vitb16_${N}gpu.json | cut -c1-200
timeout 400 $TR tools/bench_vit.py --arch vitl14 --batch 128 --steps 5 > gpurun_out/${TAG}_vitl14_${N}gpu.json 2>> gpurun_out/${TAG}_bench_${N}gpu.err; tail -1 gpurun_out/${TAG}_vitl14_${N}gpu.json | cut -c1-200
timeout 400 $TR tools/bench_train.py > gpurun_out/${TAG}_train_${N}gpu.json 2>> gpurun_out/${TAG}_bench_${N}gpu.err; tail -1 gpurun_out/${TAG}_train_${N}gpu.json | cut -c1-200
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
