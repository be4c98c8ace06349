#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_train_gpu.py -q -k "ranking or adapter or losses" 2>&1 | tail -12
timeout 600 python tools/bench_rowops.py 2>/dev/null | grep -i "rank\|kl_" | tee gpurun_out/c20_rowops.jsonl
CMD="python tools/bench_train.py --no-graph --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/c20_train_plain.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/c20_train_plain.log; exit 1; }
tail -1 gpurun_out/c20_train_plain.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file /tmp/c20_train_launches.csv $CMD > gpurun_out/c20_train_ncu.log 2>&1
echo "ncu rc=$?"
python tools/summarize_ncu.py /tmp/c20_train_launches.csv gpurun_out/c20_train_launch_summary.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/c20_train_launch_summary.json'))
print(d['launches_profiled'])
for k,v in list(d['by_kernel_instance'].items())[:45]:
    print(f"{k[:70]:70s} n={v['launches']:5d} us={v['us_total']:9.1f} share={v['share_of_profiled_time']:.3f}")
PY
