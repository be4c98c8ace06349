#!/bin/bash
# attention forward with one write_p call site (3 copies of the unrolled body instead of 9) and two row-sum chains: tests, kernel
# bench, ViT towers; then the full default bench line with its kernel table (also times the whole default run)
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_vit_gpu.py -q 2>&1 | tail -3
timeout 120 python tools/bench_attn.py 2>&1 | tail -3 | cut -c1-330
for i in 1 2; do timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('vitb16', round(d['value']), round(d['ms_per_step'],3))"; done
LECB_ATTN_POLY=0 timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('vitb16 poly0', round(d['value']), round(d['ms_per_step'],3))"
timeout 300 python tools/bench_vit.py --arch vitl14 --batch 128 --steps 5 2>&1 | grep '^{"metric' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('vitl14', round(d['value']), round(d['ms_per_step'],3))"
SECONDS=0
timeout 900 python bench.py --profile-out gpurun_out/c30_kernel_table.json > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.err; echo "bench rc=$? in ${SECONDS}s"; tail -1 gpurun_out/c30_bench.json | cut -c1-600
