#!/bin/bash
# round 2, call K (1 GPU): why does the pipelined chain not overlap? experiments with the diagnostics build
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2k}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
B="python bench.py --rows 1250000 --steps 400 --no-configs --no-cpu-baseline --recall-queries 0 --diag-lib"
echo "== diag lib, pipelined" | tee -a $L
timeout 600 $B 2>>gpurun_out/${T}.err | cut -c95-200 | tee -a $L
echo "== diag lib, pipelined, no range launch (K_a -> K_a chain)" | tee -a $L
TSC_PIPE_NO_RANGE=1 timeout 600 $B 2>>gpurun_out/${T}.err | cut -c95-200 | tee -a $L
echo "== diag lib, not pipelined" | tee -a $L
timeout 600 $B --no-pipeline 2>>gpurun_out/${T}.err | cut -c95-200 | tee -a $L
echo "== smaller CTAs (4 warps): two CTAs per SM fit with room to spare" | tee -a $L
TSC_SCAN_WARPS=4 timeout 600 $B 2>>gpurun_out/${T}.err | cut -c95-200 | tee -a $L
TSC_SCAN_WARPS=4 TSC_PIPE_NO_RANGE=1 timeout 600 $B 2>>gpurun_out/${T}.err | cut -c95-200 | tee -a $L
TSC_SCAN_WARPS=4 timeout 600 $B --no-pipeline 2>>gpurun_out/${T}.err | cut -c95-200 | tee -a $L
tail -3 gpurun_out/${T}.err | tee -a $L
