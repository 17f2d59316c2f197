#!/bin/bash
# round 2, call W (1 GPU): ncu --set full of the f16 scan at C4's shard shape (12.5M x 1536 fp16, IP, k=100)
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2w}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel -s 4 -c 1 -o gpurun_out/${T}_c4 python tools/bench_configs.py c4 > gpurun_out/${T}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_c4.ncu-rep > gpurun_out/${T}_c4_summary.txt 2>&1
ncu -i gpurun_out/${T}_c4.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${T}_c4_source.csv.gz
rm -f gpurun_out/${T}_c4.ncu-rep
cat gpurun_out/${T}_c4_summary.txt | tee -a $L
tail -2 gpurun_out/${T}_ncu.log | cut -c1-300 | tee -a $L
