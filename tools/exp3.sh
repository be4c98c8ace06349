#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_attn_gpu.py -x -q 2>&1 | tail -3
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B --profile-out gpurun_out/x5_${tag}.json 2>&1 | tail -1 | cut -c1-200; }
run base LECB_X=0
run nbh2 LECB_NB_HALO32=2
PYTHONPATH=. python tools/micro/conv_one.py 256 224 224 32 64
PYTHONPATH=. python tools/micro/conv_one.py 256 224 224 32 32
PYTHONPATH=. python tools/micro/conv_one.py 256 112 112 64 64
PYTHONPATH=. timeout 120 python tools/bench_attn.py 2>&1 | tail -6
