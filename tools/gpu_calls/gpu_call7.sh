#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
ncu --query-metrics 2>/dev/null | grep -i "tensor\|utc\|tmem" | head -60 > gpurun_out/c7_tensor_metrics.txt; wc -l gpurun_out/c7_tensor_metrics.txt
for w in stem stem_u8; do
  python tools/ncu_one.py $w > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:stem_conv1 -s 2 -c 1 -o gpurun_out/c7_$w -f python tools/ncu_one.py $w > gpurun_out/c7_ncu_$w.log 2>&1; tail -2 gpurun_out/c7_ncu_$w.log
done
for w in conv3 reduce; do
  python tools/ncu_one.py $w 1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 2 -c 1 -o gpurun_out/c7_pair_$w -f python tools/ncu_one.py $w 1 > gpurun_out/c7_ncu_pair_$w.log 2>&1; tail -2 gpurun_out/c7_ncu_pair_$w.log
  python tools/ncu_one.py $w 0 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/c7_single_$w -f python tools/ncu_one.py $w 0 > gpurun_out/c7_ncu_single_$w.log 2>&1; tail -2 gpurun_out/c7_ncu_single_$w.log
done
ls -la gpurun_out/*.ncu-rep
