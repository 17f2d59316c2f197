#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
echo "== configs" | tee gpurun_out/gemm2.log
timeout 900 python tools/bench_configs.py c1 c3 c3s c4 c5 c2b 2>&1 | tee gpurun_out/configs.jsonl | tee -a gpurun_out/gemm2.log
echo "== ncu gemm" | tee -a gpurun_out/gemm2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 1 \
  -f -o gpurun_out/prof_gemm python tools/bench_configs.py c3 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/gemm2.log
