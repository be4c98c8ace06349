#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q -k "pool or halo" 2>&1 | tail -3
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B --profile-out gpurun_out/x8_${tag}.json 2>&1 | tail -1 | cut -c1-160; }
run base LECB_X=0
run base2 LECB_X=0
