#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_pair_gemm_gpu.py -q -x -s 2>&1 > gpurun_out/c5_pair_tests.txt; tail -25 gpurun_out/c5_pair_tests.txt
if grep -q "passed" gpurun_out/c5_pair_tests.txt && ! grep -q "failed" gpurun_out/c5_pair_tests.txt; then
  for mode in 0 1; do
    LECB_NO_PAIR=$([ $mode = 0 ] && echo 1) timeout 600 env $( [ $mode = 0 ] && echo LECB_NO_PAIR=1 ) python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-gpu-reference --profile-out gpurun_out/c5_table_pair$mode.json > gpurun_out/c5_bench_pair$mode.json 2> gpurun_out/c5_bench_pair$mode.err
    python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/c5_bench_pair$mode.json') if l.startswith('{')][-1])
print('pair=$mode', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['per_layer_bound']['frac'], d['kernel_ms_per_step'])
PY
  done
fi
