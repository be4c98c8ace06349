#!/bin/bash
set -u
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B --profile-out gpurun_out/x6_${tag}.json 2>&1 | tail -1 | cut -c1-160; }
run base LECB_X=0
run a8 LECB_PF_A=8
run a16 LECB_PF_A=16
run r8 LECB_PF_R=8
run r16 LECB_PF_R=16
run a16r8 LECB_PF_A=16 LECB_PF_R=8
run a32r16 LECB_PF_A=32 LECB_PF_R=16
timeout 300 env LECB_PF_A=16 LECB_PF_R=8 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -2
