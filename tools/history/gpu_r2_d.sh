#!/bin/bash
# round 2, call D (1 GPU): trace of the reworked tail, tests, bench, WHERE kernel, sparse ncu
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2d.log
nvidia-smi -L | tee $L
echo "== scan trace: 1.25M x 768 (one of 8 shards), 10M x 768, 10k x 128" | tee -a $L
timeout 300 python tools/scan_trace.py 1250000 768 10 4 2>&1 | tail -5 | tee -a $L
timeout 300 python tools/scan_trace.py 10000000 768 10 3 2>&1 | tail -4 | tee -a $L
timeout 300 python tools/scan_trace.py 10000 128 10 3 2>&1 | tail -4 | tee -a $L
echo "== scan trace: C4 shard 12.5M x 1536 f16 IP k=100" | tee -a $L
timeout 300 python tools/scan_trace.py 12500000 1536 100 3 1 2 2>&1 | tail -4 | tee -a $L
echo "== fp64 add latency" | tee -a $L
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/dadd_latency.cu -o /tmp/dadd && /tmp/dadd 2>&1 | tee -a $L
echo "== gpu tests" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 | tee -a $L
echo "== bench" | tee -a $L
timeout 900 python bench.py --no-configs 2>gpurun_out/r2d_bench.err | tee gpurun_out/r2d_bench.json | cut -c1-1800 | tee -a $L
tail -3 gpurun_out/r2d_bench.err | tee -a $L
echo "== configs" | tee -a $L
timeout 900 python tools/bench_configs.py c1 c5w c4 2>&1 | tee gpurun_out/r2d_configs.jsonl | cut -c1-900 | tee -a $L
echo "== ncu where kernel" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:where_eval -c 3 python tools/bench_configs.py c5w 2>&1 | grep -E "gpu__time|dram__" | tee -a $L
echo "== ncu full: sparse scan (first pass)" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_sparse -s 4 -c 1 -o gpurun_out/r2d_sparse_full python tools/bench_configs.py c5 > gpurun_out/r2d_ncu_sparse.log 2>&1
python tools/ncu_summary.py gpurun_out/r2d_sparse_full.ncu-rep > gpurun_out/r2d_sparse_ncu_summary.txt 2>&1
ncu -i gpurun_out/r2d_sparse_full.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2d_sparse_source.csv.gz
rm -f gpurun_out/r2d_sparse_full.ncu-rep
head -12 gpurun_out/r2d_sparse_ncu_summary.txt | tee -a $L
