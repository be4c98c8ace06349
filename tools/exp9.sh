#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3
timeout 60 python tools/micro/gemm_one.py 200704 1024 256 res; timeout 60 python tools/micro/gemm_one.py 200704 256 1024; timeout 60 python tools/micro/gemm_one.py 3211264 256 64 res
timeout 60 python tools/micro/conv_one.py 256 224 224 32 64; timeout 60 python tools/micro/conv_one.py 256 112 112 64 64; timeout 60 python tools/micro/conv_one.py 256 28 28 256 256
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B --profile-out gpurun_out/${TAG}_${tag}.json 2>&1 | tail -1 | cut -c1-160; }
run base LECB_X=0
run base2 LECB_X=0
