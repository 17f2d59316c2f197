#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
echo "== pytest parity (all)" | tee gpurun_out/sparse.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 | tee -a gpurun_out/sparse.log
echo "== c5" | tee -a gpurun_out/sparse.log
timeout 600 python tools/bench_configs.py c5 2>&1 | tee -a gpurun_out/sparse.log
