#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/c9_tests.txt; cat gpurun_out/c9_tests.txt
timeout 600 python tools/bench_rowops.py 2>&1 | grep -i "stem\|ranking\|kl_\|crop" | tee gpurun_out/c9_rowops.jsonl
timeout 600 python tools/bench_attn.py 2>&1 | tee gpurun_out/c9_attn.jsonl
timeout 600 python tools/bench_pair_ab.py 2>&1 | tail -4 | tee gpurun_out/c9_pair_ab_f32.jsonl
