#!/bin/bash
# 2-GPU check of the final state under torchrun: default bench line (extras + sharded-prompt self-check inside) and the reference arm
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err; echo "rc=$? ${SECONDS}s"; tail -1 gpurun_out/r02f_bench_2gpu.json | cut -c1-300
SECONDS=0
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference_2gpu.json 2>> gpurun_out/r02f_bench_2gpu.err; echo "rc=$? ${SECONDS}s"; tail -1 gpurun_out/r02f_bench_reference_2gpu.json | cut -c1-200
tail -3 gpurun_out/r02f_bench_2gpu.err
