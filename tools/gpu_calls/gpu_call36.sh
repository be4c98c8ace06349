#!/bin/bash
# Stem conv1: bias through two spare K columns of the MMA (epilogue = one cvt.relu per channel pair) and the uint8 staging without
# per-byte index arithmetic on interior tiles.  gpu suite on the new build, then A/B against the previous build
# (tools/micro/liblecb_prev.so via LECB_LIB_PATH): the stem lines of the row-kernel bench and the bench step, alternating.
# Outputs: gpurun_out/c36_*
set -u
T=c36
mkdir -p gpurun_out
export PYTHONPATH=.
SECONDS=0
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? ${SECONDS}s"; tail -4 gpurun_out/${T}_pytest.log
PREV=$PWD/tools/micro/liblecb_prev.so
for i in 1 2; do
  LECB_LIB_PATH=$PREV timeout 100 python tools/bench_rowops.py 2>/dev/null | grep stem_conv1 | cut -c1-150 | sed 's/^/prev /'
  timeout 100 python tools/bench_rowops.py 2>/dev/null | grep stem_conv1 | cut -c1-150 | sed 's/^/new  /'
done | tee gpurun_out/${T}_stem_ab.txt
B="timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-gpu-reference"
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), round(d["e2e_fp32"]["value"]), d["clocks"]["sm_mhz"], d["kernel_ms_per_step"].get("lecb_stem_conv1"))'
for i in 1 2; do
  LECB_LIB_PATH=$PREV $B 2>>gpurun_out/${T}_ab.err | tail -1 | python -c "$P" prev
  $B 2>>gpurun_out/${T}_ab.err | tail -1 | python -c "$P" new
done | tee gpurun_out/${T}_bench_ab.txt
