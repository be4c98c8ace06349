#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_round2_gpu.py "tests/test_inference_gpu.py::test_forward_test_matches_reference" -q -s 2>&1 > gpurun_out/c3_tests_full.txt; tail -15 gpurun_out/c3_tests_full.txt
timeout 300 python tools/bench_retrieval.py > gpurun_out/c3_retrieval.json 2> gpurun_out/c3_retrieval.err; cat gpurun_out/c3_retrieval.json; tail -3 gpurun_out/c3_retrieval.err
