#!/bin/bash
# Round-end refresh on ONE GPU (gpurun): full gpu test suite, headline bench (+ reference arm), ncu launch list + full
# capture of the GEMM family, per-launch DRAM traffic, row-kernel roofline, secondary benches.  Outputs: gpurun_out/<TAG>_*
set -u
TAG=${1:-r01f}
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --profile-out gpurun_out/${TAG}_kernel_table.json > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -1 gpurun_out/${TAG}_bench.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; tail -1 gpurun_out/${TAG}_bench_reference.json | cut -c1-300
timeout 900 bash tools/gpu_profile.sh ${TAG} 2>&1 | tail -3
timeout 900 bash tools/gpu_profile_traffic.sh ${TAG} 2>&1 | tail -3
timeout 300 python tools/bench_vit.py --arch vitb16 --profile > gpurun_out/${TAG}_vitb16.log 2>&1; grep '^{"metric' gpurun_out/${TAG}_vitb16.log | cut -c1-250
timeout 300 python tools/bench_vit.py --arch vitl14 --batch 128 --steps 5 > gpurun_out/${TAG}_vitl14.log 2>&1; tail -1 gpurun_out/${TAG}_vitl14.log | cut -c1-250
timeout 300 python tools/bench_train.py > gpurun_out/${TAG}_train.log 2>&1; tail -1 gpurun_out/${TAG}_train.log | cut -c1-250
timeout 300 python tools/bench_train.py --graph >> gpurun_out/${TAG}_train.log 2>&1; tail -1 gpurun_out/${TAG}_train.log | cut -c1-250
timeout 120 python tools/bench_attn.py > gpurun_out/${TAG}_attn.log 2>&1; cat gpurun_out/${TAG}_attn.log
timeout 200 python tools/bench_retrieval.py > gpurun_out/${TAG}_retrieval.log 2>&1; tail -2 gpurun_out/${TAG}_retrieval.log | cut -c1-250
ncu -i gpurun_out/${TAG}_gemm_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm_full_raw.csv 2>/dev/null
python __graft_entry__.py > /dev/null && python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
