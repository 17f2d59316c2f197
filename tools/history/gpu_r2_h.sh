#!/bin/bash
# round 2, call H (1 GPU): tests (large-k tensor path), WHERE kernel after the range-test rewrite, sparse scan default
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2h
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
echo "== gpu tests" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 | tee -a $L
echo "== WHERE kernel + sparse scan" | tee -a $L
timeout 600 python tools/bench_configs.py c5w c5 2>&1 | tee gpurun_out/${T}_configs.jsonl | cut -c1-900 | tee -a $L
timeout 600 ncu --set full --clock-control none -k regex:where_eval -s 1 -c 1 -o gpurun_out/${T}_where_full python tools/bench_configs.py c5w > gpurun_out/${T}_ncu_where.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_where_full.ncu-rep > gpurun_out/${T}_where_ncu_summary.txt 2>&1
rm -f gpurun_out/${T}_where_full.ncu-rep
head -30 gpurun_out/${T}_where_ncu_summary.txt | tee -a $L
echo "== large-k batch on the tensor path: 1024 x 10M x 768 bf16 cosine k=100" | tee -a $L
timeout 600 python tools/bench_configs.py c3k 2>&1 | cut -c1-900 | tee -a $L
