#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/bench_vit.py --arch vitb16 --profile > gpurun_out/${TAG}_vitb16.log 2>&1; tail -1 gpurun_out/${TAG}_vitb16.log | cut -c1-600
timeout 300 python tools/bench_vit.py --arch vitl14 --batch 128 --steps 5 > gpurun_out/${TAG}_vitl14.log 2>&1; tail -1 gpurun_out/${TAG}_vitl14.log | cut -c1-400
timeout 300 python tools/bench_train.py > gpurun_out/${TAG}_train.log 2>&1; tail -1 gpurun_out/${TAG}_train.log | cut -c1-400
timeout 300 python tools/bench_train.py --graph >> gpurun_out/${TAG}_train.log 2>&1; tail -1 gpurun_out/${TAG}_train.log | cut -c1-400
