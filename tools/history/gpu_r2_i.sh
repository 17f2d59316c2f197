#!/bin/bash
# round 2, call I (1 GPU): pipelined device searches — tests, bench with / without, a 1.25M-row shard both ways
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2i}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
echo "== gpu tests (pipeline, where, certificate first; then everything)" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -x 2>&1 | tail -15 | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee -a $L
echo "== bench, pipelined (default)" | tee -a $L
timeout 900 python bench.py --no-configs 2>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | cut -c1-2600 | tee -a $L
tail -3 gpurun_out/${T}_bench.err | tee -a $L
echo "== bench, --no-pipeline" | tee -a $L
timeout 900 python bench.py --no-configs --no-pipeline --no-cpu-baseline --recall-queries 1 2>>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench_nopipe.json | cut -c1-700 | tee -a $L
echo "== one shard of eight (1.25M rows), both ways" | tee -a $L
timeout 600 python bench.py --rows 1250000 --steps 400 --no-configs --no-cpu-baseline --recall-queries 1 2>>gpurun_out/${T}_bench.err | cut -c1-700 | tee -a $L
timeout 600 python bench.py --rows 1250000 --steps 400 --no-configs --no-cpu-baseline --recall-queries 1 --no-pipeline 2>>gpurun_out/${T}_bench.err | cut -c1-700 | tee -a $L
echo "== WHERE kernel" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum --clock-control none -k regex:where_eval -c 2 python tools/bench_configs.py c5w 2>&1 | grep -E "gpu__time|dram__|inst_exec" | tee -a $L
