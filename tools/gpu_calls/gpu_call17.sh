#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/c17_tests.txt; cat gpurun_out/c17_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for a in "" "--no-graph" "--no-graph --no-trim" "--no-trim"; do timeout 300 python tools/bench_train.py $a 2>/dev/null | tail -1 | cut -c1-700; done | tee gpurun_out/c17_train.jsonl
timeout 1200 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/c17_kernel_table.json > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/c17_bench.json
timeout 600 python tools/bench_rowops.py > gpurun_out/c17_rowops.jsonl 2>/dev/null; wc -l gpurun_out/c17_rowops.jsonl
timeout 300 python tools/bench_retrieval.py > gpurun_out/c17_retrieval.json 2>/dev/null; cat gpurun_out/c17_retrieval.json | cut -c1-400
