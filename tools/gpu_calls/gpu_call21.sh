#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_pair_gemm_gpu.py tests/test_train_gpu.py tests/test_attn_gpu.py -q -x 2>&1 | tail -12
timeout 300 python tools/which_gemm.py 2>&1 | tail -12 | tee gpurun_out/c21_which_gemm.txt
for a in "" "--no-graph"; do timeout 300 python tools/bench_train.py $a 2>/dev/null | tail -1 | cut -c1-500; done | tee gpurun_out/c21_train.jsonl
