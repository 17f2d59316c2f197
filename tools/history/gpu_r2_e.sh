#!/bin/bash
# round 2, call E (1 GPU): trace of the tail (row staging by bulk copies, one lane per chain), tests, bench
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2e}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
echo "== scan trace: 1.25M x 768 (one of 8 shards), 10M x 768, 10k x 128" | tee -a $L
timeout 300 python tools/scan_trace.py 1250000 768 10 4 2>&1 | tail -5 | tee -a $L
timeout 300 python tools/scan_trace.py 10000000 768 10 3 2>&1 | tail -4 | tee -a $L
timeout 300 python tools/scan_trace.py 10000 128 10 3 2>&1 | tail -4 | tee -a $L
echo "== scan trace: C4 shard 12.5M x 1536 f16 IP k=100" | tee -a $L
timeout 300 python tools/scan_trace.py 12500000 1536 100 3 1 2 2>&1 | tail -4 | tee -a $L
echo "== gpu tests" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -15 | tee -a $L
echo "== bench" | tee -a $L
timeout 900 python bench.py --no-configs 2>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | cut -c1-1800 | tee -a $L
tail -3 gpurun_out/${T}_bench.err | tee -a $L
echo "== configs" | tee -a $L
timeout 900 python tools/bench_configs.py c1 c4 c3 c5 2>&1 | tee gpurun_out/${T}_configs.jsonl | cut -c1-900 | tee -a $L
