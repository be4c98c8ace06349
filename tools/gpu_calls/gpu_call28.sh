#!/bin/bash
# late PDL trigger in the persistent kernels: ViT-B / prompt-tuning / headline A/B against no PDL; attnpool rewrite + attention
# ragged-edge tests; row-kernel numbers
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_inference_gpu.py -q 2>&1 | tail -4
vit() { timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 vit  ', round(d['value']), round(d['ms_per_step'],3))"; }
trn() { timeout 200 python tools/bench_train.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 train', round(d['value']), round(d['ms_per_step'],4))"; }
hdl() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-extra 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 rn101', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'])"; }
for rep in 1 2 3; do
  LECB_PDL_MODE=0 vit "none    "
  LECB_PDL_MODE=1 vit "late-all"
  LECB_PDL_MODE=3 vit "late-hvy"
done
for rep in 1 2; do
  LECB_PDL_MODE=0 trn "none    "
  LECB_PDL_MODE=1 trn "late-all"
  LECB_PDL_MODE=1 LECB_LIB_PATH=$PWD/tools/micro/liblecb_notrig.so trn "notrig  "
done
for rep in 1 2; do
  LECB_PDL_MODE=0 hdl "none    "
  LECB_PDL_MODE=1 hdl "late-all"
done
timeout 300 python tools/bench_rowops.py 2>/dev/null | cut -c1-200 > gpurun_out/c28_rowops.jsonl; head -2 gpurun_out/c28_rowops.jsonl
timeout 120 python tools/bench_attn.py 2>&1 | tail -3 | cut -c1-200
