#!/bin/bash
# Closing call of round 2: lean ASL kernel (flush-to-zero MUFU forms, one logarithm, branch-free) — gpu suite incl. the float64
# test of all three ASL variants, A/B of the row-kernel bench against the previous build; then the refresh of the committed
# evidence on the final tree: default bench line + kernel table, ncu launch list / DRAM traffic of one step and of the row kernels.
# Outputs: gpurun_out/c37_*
set -u
T=c37
mkdir -p gpurun_out
export PYTHONPATH=.
SECONDS=0
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? ${SECONDS}s"; tail -4 gpurun_out/${T}_pytest.log
PREV=$PWD/tools/micro/liblecb_prev.so
for i in 1 2; do
  LECB_LIB_PATH=$PREV timeout 100 python tools/bench_rowops.py 2>/dev/null | grep "asl_fwd\|kl_softmax" | cut -c1-150 | sed 's/^/prev /'
  timeout 100 python tools/bench_rowops.py 2>/dev/null | grep "asl_fwd\|kl_softmax" | cut -c1-150 | sed 's/^/new  /'
done | tee gpurun_out/${T}_asl_ab.txt
timeout 100 python tools/bench_rowops.py > gpurun_out/${T}_rowops.jsonl 2>/dev/null
SECONDS=0
timeout 300 python bench.py --profile-out gpurun_out/${T}_kernel_table.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$? ${SECONDS}s"; tail -1 gpurun_out/${T}_bench.json | cut -c1-200
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-gpu-reference --ncu-window"
$CMD > gpurun_out/${T}_traffic_plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off -c 160 --csv --log-file gpurun_out/${T}_step_traffic.csv $CMD > gpurun_out/${T}_traffic_ncu.log 2>&1
echo "step traffic rc=$?"
LECB_ROWOPS_ITERS=3 timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'head_aggregate|asl_fwd_bwd|l2norm_kernel|ranking|kl_softmax|layernorm_fwd|avgpool|quick_gelu_fwd|stem_conv1|attnpool' -c 120 --csv \
    --log-file gpurun_out/${T}_rowops_traffic.csv python tools/bench_rowops.py > gpurun_out/${T}_rowops_ncu.log 2>&1
echo "rowops ncu rc=$?"
