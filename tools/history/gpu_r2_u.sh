#!/bin/bash
# round 2, call U (1 GPU): ncu of the tensor path at a small batch (8 queries, bf16): what limits it?
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2u}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
timeout 600 ncu --set full --clock-control none -k regex:gemm_topk -s 2 -c 1 -o gpurun_out/${T}_gemm8 python tools/bench_configs.py c3q > gpurun_out/${T}_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_gemm8.ncu-rep > gpurun_out/${T}_gemm8_summary.txt 2>&1
ncu -i gpurun_out/${T}_gemm8.ncu-rep --page details 2>/dev/null | grep -E "Duration|DRAM Throughput|L2 Cache Throughput|Tensor|Executed Ipc|Registers Per|Theoretical Occ|Achieved Occ|Block Limit|Mem Busy|Max Bandwidth|L1/TEX Hit|Issue Slots Busy|No Eligible" | head -40 > gpurun_out/${T}_gemm8_details.txt
rm -f gpurun_out/${T}_gemm8.ncu-rep
cat gpurun_out/${T}_gemm8_summary.txt | tee -a $L
cat gpurun_out/${T}_gemm8_details.txt | tee -a $L
tail -4 gpurun_out/${T}_ncu.log | cut -c1-400 | tee -a $L
