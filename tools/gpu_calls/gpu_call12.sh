#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1200 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/c12_kernel_table.json > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
rc=$?; echo "bench rc=$rc"; tail -c 1500 gpurun_out/c12_bench.json; grep -v "^frame\|Warning" gpurun_out/c12_bench.err | tail -8 | cut -c1-300
if [ $rc -ne 0 ]; then
  CUDA_LAUNCH_BLOCKING=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-gpu-reference > gpurun_out/c12_bench_blocking.json 2> gpurun_out/c12_bench_blocking.err
  echo "blocking rc=$?"; grep -v "^frame\|Warning" gpurun_out/c12_bench_blocking.err | tail -12 | cut -c1-400
fi
timeout 600 python -m pytest tests/test_round2_gpu.py -q -x -k "window_pipeline or resample" 2>&1 | tail -4
nvidia-smi --query-gpu=name,temperature.gpu,power.draw --format=csv
dmesg 2>/dev/null | grep -i xid | tail -3
