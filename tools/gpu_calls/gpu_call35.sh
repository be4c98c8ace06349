#!/bin/bash
# Projection-shortcut tail of the four stage-entry bottlenecks as ONE K-concatenated GEMM (lecb_gemm_bf16_dual): kernel tests,
# the whole gpu suite, A/B of the bench step against the two-GEMM form (LECB_NO_DUAL=1, same library, same box, alternating),
# then the default bench line and the ncu launch list + DRAM traffic of one step of the new launch sequence.
# Outputs: gpurun_out/c35_*
set -u
T=c35
mkdir -p gpurun_out
export PYTHONPATH=.
SECONDS=0
timeout 200 python -m pytest tests/test_pair_gemm_gpu.py -q -x -k "dual" > gpurun_out/${T}_pytest_dual.log 2>&1; echo "dual tests rc=$? ${SECONDS}s"; tail -4 gpurun_out/${T}_pytest_dual.log
SECONDS=0
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? ${SECONDS}s"; tail -4 gpurun_out/${T}_pytest.log
B="timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-gpu-reference"
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["roofline"]["frac"], d["roofline"]["per_layer_bound"]["frac"], d["gpu_launches"])'
for i in 1 2; do
  LECB_NO_DUAL=1 $B --profile-out gpurun_out/${T}_two$i.json 2>>gpurun_out/${T}_ab.err | tail -1 | python -c "$P" two-gemm
  $B --profile-out gpurun_out/${T}_dual$i.json 2>>gpurun_out/${T}_ab.err | tail -1 | python -c "$P" dual
done
SECONDS=0
timeout 300 python bench.py --profile-out gpurun_out/${T}_kernel_table.json > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$? ${SECONDS}s"; tail -1 gpurun_out/${T}_bench.json | cut -c1-200
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra --no-gpu-reference --ncu-window"
$CMD > gpurun_out/${T}_traffic_plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off -c 160 --csv --log-file gpurun_out/${T}_step_traffic.csv $CMD > gpurun_out/${T}_traffic_ncu.log 2>&1
echo "step traffic rc=$?"
python tools/summarize_ncu.py gpurun_out/${T}_step_traffic.csv gpurun_out/${T}_step_traffic_summary.json | head -6
