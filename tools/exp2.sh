#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "halo or conv3x3" 2>&1 | tail -3
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B --profile-out gpurun_out/x3_${tag}.json 2>&1 | tail -1 | cut -c1-200; }
run base LECB_X=0
run nbh2 LECB_NB_HALO32=2
PYTHONPATH=. python tools/micro/conv_one.py 256 224 224 32 64
PYTHONPATH=. python tools/micro/conv_one.py 256 224 224 32 32
PYTHONPATH=. python tools/micro/conv_one.py 256 112 112 64 64
