#!/bin/bash
# head_aggregate rewrite (fixed-reference fast path + online fallback): kernel tests, model-level tests, row-kernel roofline;
# PDL A/B on the ViT-B tower, three alternations
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_head_gpu.py -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/bench_rowops.py 2>/dev/null | head -3 | cut -c1-200
for i in 1 2 3; do
  timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pdl  ', round(d['value']), d['ms_per_step'])"
  LECB_NO_PDL=1 timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nopdl', round(d['value']), d['ms_per_step'])"
done
