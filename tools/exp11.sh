#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -k "descending or resident" 2>&1 | tail -2
B="timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline"
for i in 1 2 3; do
  LECB_NO_DESCENDING=1 $B --profile-out gpurun_out/${TAG}_off$i.json 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('off', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'])"
  $B --profile-out gpurun_out/${TAG}_on$i.json 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('on ', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'])"
done
