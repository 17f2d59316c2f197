#!/bin/bash
set +e
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r1b_3.log
echo "== pytest gemm" | tee $L
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -5 | tee -a $L
echo "== matrix" | tee -a $L
timeout 900 python tools/gemm_matrix.py "PROF=1" "EXP=2,PROF=1" "STAGES=4,PROF=1" "2CTA=0,PROF=1" "" "" 2>&1 | tee -a $L
echo "== ncu metrics" | tee -a $L
timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second,l1tex__m_xbar2l1tex_read_bytes.sum \
  --clock-control none -k regex:gemm_topk -s 2 -c 1 python tools/bench_configs.py c3 2>&1 | grep -E "gemm_topk|dram__|gpu__time|tensor|hit_rate|per_second|xbar2l1tex" | tee -a $L
