#!/bin/bash
# ncu captures, summarised ON THE BOX (the reports of the many-instantiation kernel are 30 MB each; gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out /tmp/ncu
python -c "import __graft_entry__ as g; g.build()" || exit 1
ncu --query-metrics 2>/dev/null | grep -i "tensor\|utc\|tmem" | head -80 > gpurun_out/c8_tensor_metrics.txt; wc -l gpurun_out/c8_tensor_metrics.txt
cap() {   # name, kernel regex, launcher args...
  local name=$1 rx=$2; shift 2
  python tools/ncu_one.py "$@" > /dev/null 2>&1 || { echo "plain run failed: $name"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -o /tmp/ncu/$name -f python tools/ncu_one.py "$@" > /tmp/ncu/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/c8_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/$name.ncu-rep --page details > gpurun_out/c8_${name}_details.txt 2>/dev/null
  python tools/ncu_waits.py /tmp/ncu/$name.ncu-rep > gpurun_out/c8_${name}_waits.txt 2>&1
  ls -la /tmp/ncu/$name.ncu-rep | awk '{print $5, $9}'
}
cap stem stem_conv1 stem
cap stem_u8 stem_conv1 stem_u8
cap pair_conv3 gemm_pair conv3 1
cap single_conv3 gemm_kernel conv3 0
cap pair_reduce gemm_pair reduce 1
cap single_reduce gemm_kernel reduce 0
cap pair_expand gemm_pair expand 1
cp /tmp/ncu/stem.ncu-rep /tmp/ncu/pair_conv3.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
