#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_pair_gemm_gpu.py -q -x 2>&1 | tail -3
for i in 1 2 3 4 5 6 7 8; do
  timeout 300 python bench.py --steps 150 --warmup 3 --no-extra --no-cpu-baseline --no-gpu-reference > gpurun_out/c16_stress_$i.json 2> gpurun_out/c16_stress_$i.err
  echo "run $i rc=$? $(grep -c 'launch failure' gpurun_out/c16_stress_$i.err) $(grep -h 'timed out' gpurun_out/c16_stress_$i.json gpurun_out/c16_stress_$i.err | head -1)"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/c16_stress_*.json')):
    l=[x for x in open(f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
PY
