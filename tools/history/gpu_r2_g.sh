#!/bin/bash
# round 2, call G (1 GPU): tests, bench, WHERE kernel (4 words per step), sparse scan at 16 warps,
# ncu --set full of the tcgen05 GEMM (C3) and of the dense scan, launch list of the bench command
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2g
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
echo "== scan trace: 1.25M x 768, 10k x 128" | tee -a $L
timeout 300 python tools/scan_trace.py 1250000 768 10 4 2>&1 | tail -3 | tee -a $L
timeout 300 python tools/scan_trace.py 10000 128 10 3 2>&1 | tail -2 | tee -a $L
echo "== gpu tests" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -8 | tee -a $L
echo "== bench (full line)" | tee -a $L
timeout 900 python bench.py 2>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | cut -c1-1500 | tee -a $L
tail -3 gpurun_out/${T}_bench.err | tee -a $L
echo "== WHERE kernel + sparse scan" | tee -a $L
timeout 600 python tools/bench_configs.py c5w c5 2>&1 | tee gpurun_out/${T}_configs.jsonl | cut -c1-900 | tee -a $L
for w in 8 16; do
  echo "-- sparse scan, TSC_SCAN_WARPS=$w (diagnostics build)" | tee -a $L
  TSC_SCAN_WARPS=$w timeout 600 python tools/bench_configs.py --diag c5 2>&1 | cut -c1-500 | tee -a $L
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:where_eval -c 3 python tools/bench_configs.py c5w 2>&1 | grep -E "gpu__time|dram__" | tee -a $L
echo "== ncu full: tcgen05 GEMM (C3)" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 3 -c 1 -o gpurun_out/${T}_gemm_full python tools/bench_configs.py c3 > gpurun_out/${T}_ncu_gemm.log 2>&1
tail -2 gpurun_out/${T}_ncu_gemm.log | cut -c1-300 | tee -a $L
python tools/ncu_summary.py gpurun_out/${T}_gemm_full.ncu-rep > gpurun_out/${T}_gemm_ncu_summary.txt 2>&1
ncu -i gpurun_out/${T}_gemm_full.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${T}_gemm_source.csv.gz
ncu -i gpurun_out/${T}_gemm_full.ncu-rep --page details 2>/dev/null | head -400 > gpurun_out/${T}_gemm_details.txt
rm -f gpurun_out/${T}_gemm_full.ncu-rep
head -30 gpurun_out/${T}_gemm_ncu_summary.txt | tee -a $L
echo "== ncu full: dense scan (headline)" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel -s 6 -c 1 -o gpurun_out/${T}_scan_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --recall-queries 0 > gpurun_out/${T}_ncu_scan.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_scan_full.ncu-rep > gpurun_out/${T}_scan_ncu_summary.txt 2>&1
rm -f gpurun_out/${T}_scan_full.ncu-rep
head -12 gpurun_out/${T}_scan_ncu_summary.txt | tee -a $L
echo "== launch list of the bench command" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs --recall-queries 0 > gpurun_out/${T}_ncu_bench.log 2>&1
tail -1 gpurun_out/${T}_ncu_bench.log | cut -c1-300 | tee -a $L
