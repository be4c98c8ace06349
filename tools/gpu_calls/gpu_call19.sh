#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_round2_gpu.py -q -k "adapter or window_pipeline" 2>&1 | tail -30
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/c19_tests.txt; cat gpurun_out/c19_tests.txt
