#!/bin/bash
# round 2, call C (1 GPU): where the scan launch spends its time (diagnostics library), the
# full bench line with the configs block, ncu captures of the dense and sparse scan kernels.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2c.log
nvidia-smi -L | tee $L
echo "== scan trace: 1.25M x 768 (one of 8 shards), 10M x 768" | tee -a $L
timeout 300 python tools/scan_trace.py 1250000 768 10 4 2>&1 | tail -5 | tee -a $L
timeout 300 python tools/scan_trace.py 10000000 768 10 3 2>&1 | tail -4 | tee -a $L
timeout 300 python tools/scan_trace.py 10000 128 10 3 2>&1 | tail -4 | tee -a $L
echo "== gpu tests" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee -a $L
echo "== bench (full line)" | tee -a $L
timeout 900 python bench.py 2>gpurun_out/r2c_bench.err | tee gpurun_out/r2c_bench.json | cut -c1-3000 | tee -a $L
tail -3 gpurun_out/r2c_bench.err | tee -a $L
echo "== ncu full: dense scan" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel -s 6 -c 1 -o gpurun_out/r2c_scan_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --recall-queries 0 > gpurun_out/r2c_ncu_scan.log 2>&1
tail -2 gpurun_out/r2c_ncu_scan.log | tee -a $L
python tools/ncu_summary.py gpurun_out/r2c_scan_full.ncu-rep > gpurun_out/r2c_scan_ncu_summary.txt 2>&1
ncu -i gpurun_out/r2c_scan_full.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2c_scan_source.csv.gz
rm -f gpurun_out/r2c_scan_full.ncu-rep
echo "== ncu full: sparse scan" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_sparse -s 3 -c 1 -o gpurun_out/r2c_sparse_full python tools/bench_configs.py c5 > gpurun_out/r2c_ncu_sparse.log 2>&1
tail -2 gpurun_out/r2c_ncu_sparse.log | tee -a $L
python tools/ncu_summary.py gpurun_out/r2c_sparse_full.ncu-rep > gpurun_out/r2c_sparse_ncu_summary.txt 2>&1
ncu -i gpurun_out/r2c_sparse_full.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2c_sparse_source.csv.gz
rm -f gpurun_out/r2c_sparse_full.ncu-rep
du -sh gpurun_out | tee -a $L
echo "== launch list of the bench command" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs --recall-queries 0 > gpurun_out/r2c_ncu_bench.log 2>&1
tail -1 gpurun_out/r2c_ncu_bench.log | cut -c1-300 | tee -a $L
