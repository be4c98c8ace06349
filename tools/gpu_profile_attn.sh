#!/bin/bash
# ncu --set full on the tcgen05 attention kernel at the ViT-B/16 shape.
set -u
mkdir -p gpurun_out
LECB_DEBUG=1 python tools/bench_attn.py > gpurun_out/attn_plain.log 2>&1 || { tail -5 gpurun_out/attn_plain.log; exit 1; }
cat gpurun_out/attn_plain.log
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 3 -c 1 \
    -o gpurun_out/${1:-r01}_attn_full python tools/bench_attn.py > gpurun_out/attn_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/attn_ncu.log
ncu -i gpurun_out/${1:-r01}_attn_full.ncu-rep --page raw --csv > gpurun_out/${1:-r01}_attn_raw.csv 2>/dev/null
