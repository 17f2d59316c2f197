#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
echo "== gemm_dev small" | tee gpurun_out/gemm1.log
timeout 120 python tools/gemm_dev.py 2>&1 | tail -12 | tee -a gpurun_out/gemm1.log
echo "== gemm_dev 768" | tee -a gpurun_out/gemm1.log
D=768 NR=1000 NQ=200 timeout 120 python tools/gemm_dev.py 2>&1 | tail -12 | tee -a gpurun_out/gemm1.log
echo "== pytest gemm" | tee -a gpurun_out/gemm1.log
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -25 | tee -a gpurun_out/gemm1.log
