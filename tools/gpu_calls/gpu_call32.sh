#!/bin/bash
# retrieval call time under the PDL modes (the per-call time doubled against the first half of the round while the kernel times did not)
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
for mode in 0 1 3 0 1; do
  LECB_PDL_MODE=$mode timeout 200 python tools/bench_retrieval.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode$mode', d['ms'], d['kernels_ms'])"
done
