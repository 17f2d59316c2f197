#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
echo "== pytest gemm" | tee gpurun_out/gemm7.log
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -5 | tee -a gpurun_out/gemm7.log
echo "== c3 variants: EXP L2PF" | tee -a gpurun_out/gemm7.log
for cfg in "0 0" "0 1" "0 2" "0 4" "2 1" "2 2" "8 2" "10 2"; do
set -- $cfg
TSC_GEMM_EXP=$1 TSC_GEMM_L2PF=$2 timeout 300 python tools/bench_configs.py c3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('exp=$1 l2pf=$2', 'hot_ms=%.2f'%d['hot_kernel_ms'], 'TF=%.0f'%d.get('tflops',0), 'frac_burst=%.3f'%d.get('tensor_frac_of_measured_burst',0), 'total_ms=%.2f'%d['device_ms_per_search'])
" | tee -a gpurun_out/gemm7.log
done
echo "== ncu gemm" | tee -a gpurun_out/gemm7.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 1 \
  -f -o gpurun_out/prof_gemm python tools/bench_configs.py c3 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/gemm7.log
