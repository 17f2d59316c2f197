#!/bin/bash
# round 2, call A: the rebuilt search pipeline (fused tail, certificate, range pass) on 1 GPU
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2a.log
nvidia-smi -L | tee $L
echo "== new tests first" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_certificate.py -m gpu -q --timeout 300 -x 2>&1 | tail -40 | tee -a $L
echo "== full gpu suite" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 | tee -a $L
echo "== bench" | tee -a $L
timeout 600 python bench.py --steps 50 2>gpurun_out/r2a_bench.err | tee gpurun_out/r2a_bench.json | tee -a $L
tail -5 gpurun_out/r2a_bench.err | tee -a $L
echo "== configs" | tee -a $L
timeout 900 python tools/bench_configs.py c1 c3 c5 c2t 2>&1 | tee gpurun_out/r2a_configs.jsonl | tee -a $L
