#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_train_gpu.py tests/test_attn_gpu.py tests/test_vit_gpu.py -q -x 2>&1 | tail -6 > gpurun_out/c10_tests.txt; cat gpurun_out/c10_tests.txt
timeout 600 python tools/bench_rowops.py 2>&1 | grep -i "stem" | tee gpurun_out/c10_rowops.jsonl
timeout 1200 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/c10_kernel_table.json > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; tail -c 2500 gpurun_out/c10_bench.json; tail -3 gpurun_out/c10_bench.err
