#!/bin/bash
# A/B of two library builds on the same box: tools/micro/liblecb_prev.so vs the in-tree build, alternating
set -u
mkdir -p gpurun_out
B="timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline"
for i in 1 2 3; do
  LECB_LIB_PATH=$PWD/tools/micro/liblecb_prev.so $B --profile-out gpurun_out/${TAG}_prev$i.json 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prev', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'])"
  $B --profile-out gpurun_out/${TAG}_new$i.json 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('new ', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'])"
done
