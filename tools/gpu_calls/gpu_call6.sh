#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python tools/bench_pair_ab.py > gpurun_out/c6_pair_ab.jsonl 2> gpurun_out/c6_pair_ab.err; cat gpurun_out/c6_pair_ab.jsonl; tail -3 gpurun_out/c6_pair_ab.err
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_round2_gpu.py tests/test_inference_gpu.py tests/test_fullsize_gpu.py -q 2>&1 | tail -8 > gpurun_out/c6_tests.txt; cat gpurun_out/c6_tests.txt
timeout 600 python tools/bench_rowops.py 2>&1 | grep -i "stem\|ranking\|kl_\|crop" | tee gpurun_out/c6_rowops.jsonl
