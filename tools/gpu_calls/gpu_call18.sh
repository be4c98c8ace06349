#!/bin/bash
# 8-GPU run of the driver's launch line (product arm with extras)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c18_bench_8gpu.json 2> gpurun_out/c18_bench_8gpu.err
echo "rc=$?"; tail -c 1200 gpurun_out/c18_bench_8gpu.json; grep -v "^frame\|Warning\|warn\|Consider\|final_loss\|run_backward" gpurun_out/c18_bench_8gpu.err | tail -5 | cut -c1-300
