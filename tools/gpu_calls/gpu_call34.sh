#!/bin/bash
# Last GPU call of round 2 (ONE GPU, ~10 min): the gpu test suite and smoke() on the final tree (the library was rebuilt in a
# re-created build container), the default bench line with its wall time, and the ncu DRAM-traffic pass of the row-kernel
# roofline bench (the round-2 rewrites of head_aggregate / ranking / KL / attnpool had CUDA-event numbers only).
# Outputs: gpurun_out/c34_*
set -u
T=c34
mkdir -p gpurun_out
export PYTHONPATH=.
SECONDS=0
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/${T}_pytest.log
SECONDS=0
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; echo "smoke ${SECONDS}s"
SECONDS=0
timeout 100 python tools/bench_rowops.py > gpurun_out/${T}_rowops.jsonl 2> gpurun_out/${T}_rowops.err; echo "rowops rc=$? ${SECONDS}s"; cut -c1-170 gpurun_out/${T}_rowops.jsonl
SECONDS=0
LECB_ROWOPS_ITERS=3 timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'head_aggregate|asl_fwd_bwd|l2norm_kernel|ranking|kl_softmax|layernorm_fwd|avgpool|quick_gelu_fwd|stem_conv1|attnpool' -c 120 --csv \
    --log-file gpurun_out/${T}_rowops_traffic.csv python tools/bench_rowops.py > gpurun_out/${T}_rowops_ncu.log 2>&1
echo "rowops ncu rc=$? ${SECONDS}s"
python tools/summarize_ncu.py gpurun_out/${T}_rowops_traffic.csv gpurun_out/${T}_rowops_traffic_summary.json | tail -14
SECONDS=0
timeout 420 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$? ${SECONDS}s"; tail -1 gpurun_out/${T}_bench.json | cut -c1-260
