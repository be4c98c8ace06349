#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 > gpurun_out/c4_tests_full.txt; tail -12 gpurun_out/c4_tests_full.txt
timeout 600 python tools/bench_rowops.py > gpurun_out/c4_rowops.jsonl 2> gpurun_out/c4_rowops.err; grep -i "stem\|ranking\|kl_\|crop" gpurun_out/c4_rowops.jsonl; tail -3 gpurun_out/c4_rowops.err
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/c4_kernel_table.json > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; tail -c 3000 gpurun_out/c4_bench.json; tail -5 gpurun_out/c4_bench.err
