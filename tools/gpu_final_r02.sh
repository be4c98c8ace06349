#!/bin/bash
# Round-2 final refresh on ONE GPU (gpurun): gpu test suite, smoke, the default bench line (+ kernel table) and the reference arm,
# ncu launch list + DRAM traffic of one steady-state step of the same command, row-kernel roofline, prompt-tuning step (graph and
# eager), attention and retrieval benches.  Outputs: gpurun_out/r02f_*
set -u
T=r02f
mkdir -p gpurun_out
export PYTHONPATH=.
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
SECONDS=0
timeout 900 python bench.py --profile-out gpurun_out/${T}_kernel_table.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$? ${SECONDS}s"; tail -1 gpurun_out/${T}_bench.json | cut -c1-200
SECONDS=0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; echo "reference rc=$? ${SECONDS}s"; tail -1 gpurun_out/${T}_bench_reference.json | cut -c1-200
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-gpu-reference --ncu-window"
$CMD > gpurun_out/${T}_traffic_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off -c 160 --csv --log-file gpurun_out/${T}_step_traffic.csv $CMD > gpurun_out/${T}_traffic_ncu.log 2>&1
echo "step traffic rc=$?"
python tools/summarize_ncu.py gpurun_out/${T}_step_traffic.csv gpurun_out/${T}_step_traffic_summary.json | tail -3
timeout 300 python tools/bench_rowops.py > gpurun_out/${T}_rowops.jsonl 2> gpurun_out/${T}_rowops.err; cut -c1-170 gpurun_out/${T}_rowops.jsonl
timeout 300 python tools/bench_train.py > gpurun_out/${T}_train.jsonl 2> gpurun_out/${T}_train.err; timeout 300 python tools/bench_train.py --no-graph >> gpurun_out/${T}_train.jsonl 2>> gpurun_out/${T}_train.err; cut -c1-120 gpurun_out/${T}_train.jsonl
timeout 120 python tools/bench_attn.py > gpurun_out/${T}_attn.jsonl 2>&1; cut -c1-230 gpurun_out/${T}_attn.jsonl
timeout 200 python tools/bench_retrieval.py > gpurun_out/${T}_retrieval.log 2>&1; tail -1 gpurun_out/${T}_retrieval.log | cut -c1-300
