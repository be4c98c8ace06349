#!/bin/bash
# adaptive PDL policy (mode 1) against no PDL (mode 0) on the three step shapes; full gpu test suite; one ncu --set full capture
# of the attention forward kernel with its wait sites
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
vit() { timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 vit  ', round(d['value']), round(d['ms_per_step'],3))"; }
trn() { timeout 200 python tools/bench_train.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 train', round(d['value']), round(d['ms_per_step'],4))"; }
hdl() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-extra 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 rn101', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['clocks']['sm_mhz'])"; }
for rep in 1 2 3; do
  LECB_PDL_MODE=0 vit "none    "
  LECB_PDL_MODE=1 vit "adaptive"
done
for rep in 1 2; do
  LECB_PDL_MODE=0 trn "none    "
  LECB_PDL_MODE=1 trn "adaptive"
  LECB_PDL_MODE=4 trn "all     "
done
for rep in 1 2 3; do
  LECB_PDL_MODE=0 hdl "none    "
  LECB_PDL_MODE=1 hdl "adaptive"
done
timeout 600 bash tools/gpu_profile_attn.sh r02b > gpurun_out/c29_attn_profile.log 2>&1; tail -2 gpurun_out/c29_attn_profile.log
timeout 200 python tools/ncu_waits.py gpurun_out/r02b_attn_full.ncu-rep > gpurun_out/r02b_attn_waits.txt 2>&1; head -40 gpurun_out/r02b_attn_waits.txt
