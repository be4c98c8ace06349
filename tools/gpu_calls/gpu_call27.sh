#!/bin/bash
# PDL policy experiment on the ViT-B tower and the prompt-tuning step: attribute on every kernel / row kernels only / GEMM +
# attention only / none, with and without the early launch_dependents trigger (second build); attention ragged-edge skip tests
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_vit_gpu.py -q 2>&1 | tail -4
timeout 120 python tools/bench_attn.py 2>&1 | tail -4
vit() { timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 vit  ', round(d['value']), round(d['ms_per_step'],3))"; }
trn() { timeout 200 python tools/bench_train.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 train', round(d['value']), round(d['ms_per_step'],4))"; }
for rep in 1 2; do
  for mode in 0 1 2 3; do
    LECB_PDL_MODE=$mode vit "trig mode$mode"
    LECB_PDL_MODE=$mode LECB_LIB_PATH=$PWD/tools/micro/liblecb_notrig.so vit "notrig mode$mode"
  done
done
for mode in 0 1 2 3; do
  LECB_PDL_MODE=$mode trn "trig mode$mode"
  LECB_PDL_MODE=$mode LECB_LIB_PATH=$PWD/tools/micro/liblecb_notrig.so trn "notrig mode$mode"
done
