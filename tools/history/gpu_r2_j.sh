#!/bin/bash
# round 2, call J (1 GPU): pipelining with the max-shared carveout
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2j}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -x 2>&1 | tail -5 | tee -a $L
echo "== one shard of eight (1.25M rows): pipelined, then not" | tee -a $L
timeout 600 python bench.py --rows 1250000 --steps 400 --no-configs --no-cpu-baseline --recall-queries 1 2>>gpurun_out/${T}_bench.err | cut -c1-330 | tee -a $L
timeout 600 python bench.py --rows 1250000 --steps 400 --no-configs --no-cpu-baseline --recall-queries 1 --no-pipeline 2>>gpurun_out/${T}_bench.err | cut -c1-330 | tee -a $L
echo "== 10M rows: pipelined, then not" | tee -a $L
timeout 900 python bench.py --no-configs --no-cpu-baseline --recall-queries 1 2>>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | cut -c1-330 | tee -a $L
timeout 900 python bench.py --no-configs --no-cpu-baseline --recall-queries 1 --no-pipeline 2>>gpurun_out/${T}_bench.err | cut -c1-330 | tee -a $L
tail -3 gpurun_out/${T}_bench.err | tee -a $L
