#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/c23_tests.txt; cat gpurun_out/c23_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
