#!/bin/bash
# PDL (programmatic dependent launch) validation: gpu tests with PDL twice + once without, then A/B of the prompt-tuning
# step and the headline step in separate processes on the same box
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
for i in 1 2; do timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/c25_tests_pdl$i.txt; tail -3 gpurun_out/c25_tests_pdl$i.txt; done
LECB_NO_PDL=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/c25_tests_nopdl.txt; tail -3 gpurun_out/c25_tests_nopdl.txt
for i in 1 2; do
  timeout 200 python tools/bench_train.py 2>&1 | tail -1 | cut -c1-120 | sed 's/^/pdl   /'
  LECB_NO_PDL=1 timeout 200 python tools/bench_train.py 2>&1 | tail -1 | cut -c1-120 | sed 's/^/nopdl /'
done
timeout 200 python tools/bench_train.py --no-graph 2>&1 | tail -1 | cut -c1-120 | sed 's/^/pdl eager   /'
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-extra"
for i in 1 2; do
  $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pdl  ', round(d['value']), d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"
  LECB_NO_PDL=1 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nopdl', round(d['value']), d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"
done
timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | cut -c1-160 | sed 's/^/pdl   /'
LECB_NO_PDL=1 timeout 200 python tools/bench_vit.py --arch vitb16 2>&1 | grep '^{"metric' | cut -c1-160 | sed 's/^/nopdl /'
