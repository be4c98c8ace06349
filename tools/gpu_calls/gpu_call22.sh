#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1200 python -m pytest tests/test_train_gpu.py tests/test_round2_gpu.py tests/test_dropin_reference_host.py -q -x 2>&1 | tail -12
for a in "" "--no-graph"; do timeout 300 python tools/bench_train.py $a 2>/dev/null | tail -1 | cut -c1-700; done | tee gpurun_out/c22_train.jsonl
