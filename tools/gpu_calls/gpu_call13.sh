#!/bin/bash
# 2-GPU validation of the driver's launch line: product arm (extras + sharded-prompt self-check) and the reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c13_bench_2gpu.json 2> gpurun_out/c13_bench_2gpu.err
echo "rc=$?"; tail -c 1800 gpurun_out/c13_bench_2gpu.json; grep -v "^frame\|Warning\|warn" gpurun_out/c13_bench_2gpu.err | tail -6 | cut -c1-300
