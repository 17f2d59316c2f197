#!/bin/bash
# session 2, call 1: re-verify the rebuilt library, then the 2-CTA isolation matrix
set +e
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r1b_1.log
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee $L
echo "== pytest gemm + loader" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_ngh_loader.py -m gpu -x -q 2>&1 | tail -5 | tee -a $L
echo "== matrix" | tee -a $L
timeout 900 python tools/gemm_matrix.py "PROF=1" "DEPTH=1,PROF=1" "DEPTH=2,PROF=1" "DEPTH=3,PROF=1" \
  "EXP=2,PROF=1" "EXP=16,PROF=1" "EXP=32,PROF=1" "EXP=1,PROF=1" "EXP=8,PROF=1" "EXP=9,PROF=1" "EXP=4,PROF=1" "EXP=6,PROF=1" "EXP=3,PROF=1" \
  "2CTA=0,PROF=1" "2CTA=0,DEPTH=2,PROF=1" "2CTA=0,EXP=2,PROF=1" "STAGES=4,DEPTH=2,PROF=1" "STAGES=3,DEPTH=2" "" 2>&1 | tee -a $L
