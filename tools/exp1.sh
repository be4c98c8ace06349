#!/bin/bash
# experiment driver (gpurun): Cin=32 halo conv tests + per-shape step profiles under dispatch knobs
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "halo or conv3x3" 2>&1 | tail -5
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" $B --profile-out gpurun_out/x1_${tag}.json 2>&1 | tail -1 | cut -c1-200; }
run base LECB_X=0
run pairs LECB_STEM_PAIRS=1
run nbh2 LECB_NB_HALO32=2
run bn128 LECB_EXPAND_BN128=1
run ring2 LECB_RING_MIN=2 LECB_NB_K4=4
run k1res LECB_RES_K1=1
run k2nb6 LECB_NB_K2=6
