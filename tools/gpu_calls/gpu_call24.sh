#!/bin/bash
# state check after the container was re-created: gpu tests, smoke, prompt-tuning + row-kernel numbers of HEAD
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/c24_tests.txt; cat gpurun_out/c24_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python tools/bench_train.py --graph > gpurun_out/c24_train.log 2>&1; tail -1 gpurun_out/c24_train.log | cut -c1-400
timeout 300 python tools/bench_rowops.py > gpurun_out/c24_rowops.jsonl 2> gpurun_out/c24_rowops.err; cut -c1-200 gpurun_out/c24_rowops.jsonl
