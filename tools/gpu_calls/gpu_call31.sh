#!/bin/bash
# attention forward: dead-warp skipping only (main build) against dead-warp + half-block variant (prev = commit 2a0f61f),
# alternating, kernel bench + ViT-B step with both polynomial settings
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 300 python -m pytest tests/test_attn_gpu.py -q 2>&1 | tail -2
for i in 1 2; do
  echo "main:"; timeout 120 python tools/bench_attn.py 2>&1 | tail -3 | cut -c1-215
  echo "prev:"; LECB_LIB_PATH=$PWD/tools/micro/liblecb_prev.so timeout 120 python tools/bench_attn.py 2>&1 | tail -3 | cut -c1-215
done
vit() { timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],3))"; }
for i in 1 2; do
  LECB_ATTN_POLY=0 vit "main poly0"
  LECB_ATTN_POLY=4 vit "main poly4"
  LECB_ATTN_POLY=0 LECB_LIB_PATH=$PWD/tools/micro/liblecb_prev.so vit "prev poly0"
  LECB_ATTN_POLY=4 LECB_LIB_PATH=$PWD/tools/micro/liblecb_prev.so vit "prev poly4"
done
