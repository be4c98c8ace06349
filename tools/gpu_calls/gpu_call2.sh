#!/bin/bash
# round-2 call 2: new tests first, then the full GPU suite, then quick benches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "== new tests"; timeout 900 python -m pytest tests/test_round2_gpu.py -q -x -s 2>&1 | tail -60 > gpurun_out/c2_new_tests.txt; tail -40 gpurun_out/c2_new_tests.txt
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/c2_full_tests.txt; tail -15 gpurun_out/c2_full_tests.txt
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/c2_kernel_table.json > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; tail -c 1500 gpurun_out/c2_bench.json; tail -3 gpurun_out/c2_bench.err
echo "== rowops"; timeout 600 python tools/bench_rowops.py > gpurun_out/c2_rowops.jsonl 2> gpurun_out/c2_rowops.err; cat gpurun_out/c2_rowops.jsonl; tail -3 gpurun_out/c2_rowops.err
echo "== retrieval"; timeout 300 python tools/bench_retrieval.py > gpurun_out/c2_retrieval.json 2> gpurun_out/c2_retrieval.err; cat gpurun_out/c2_retrieval.json; tail -3 gpurun_out/c2_retrieval.err
