#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3
python tools/micro/gemm_one.py 200704 1024 256 res; python tools/micro/gemm_one.py 200704 256 1024; python tools/micro/gemm_one.py 3211264 256 64 res
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B --profile-out gpurun_out/${TAG}_${tag}.json 2>&1 | tail -1 | cut -c1-160; }
run base LECB_X=0
run base2 LECB_X=0
