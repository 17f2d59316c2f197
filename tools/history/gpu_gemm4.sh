#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
echo "== gemm_dev (TS)" | tee gpurun_out/gemm4.log
D=768 NR=1000 NQ=200 timeout 120 python tools/gemm_dev.py 2>&1 | tail -8 | tee -a gpurun_out/gemm4.log
echo "== pytest gemm" | tee -a gpurun_out/gemm4.log
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -12 | tee -a gpurun_out/gemm4.log
echo "== c3" | tee -a gpurun_out/gemm4.log
for cfg in "1 10" "1 12" "1 8" "0 3"; do
set -- $cfg
TSC_GEMM_TS=$1 TSC_GEMM_STAGES=$2 timeout 600 python tools/bench_configs.py c3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l); continue
    print('ts=$1 stages=$2', 'hot_ms=%.2f'%d['hot_kernel_ms'], 'TF=%.0f'%d.get('tflops',0), 'frac_burst=%.3f'%d.get('tensor_frac_of_measured_burst',0), 'total_ms=%.2f'%d['device_ms_per_search'])
" | tee -a gpurun_out/gemm4.log
done
echo "== ncu gemm" | tee -a gpurun_out/gemm4.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 1 \
  -f -o gpurun_out/prof_gemm python tools/bench_configs.py c3 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/gemm4.log
