#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
echo "== gemm_dev 2cta (timeout 60)" | tee gpurun_out/gemm10.log
D=768 NR=1000 NQ=300 timeout 60 python tools/gemm_dev.py 2>&1 | tail -8 | tee -a gpurun_out/gemm10.log
echo "rc=$?" | tee -a gpurun_out/gemm10.log
echo "== pytest gemm (2cta default)" | tee -a gpurun_out/gemm10.log
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -8 | tee -a gpurun_out/gemm10.log
echo "== c3: 2CTA STAGES" | tee -a gpurun_out/gemm10.log
for cfg in "1 5" "1 4" "1 3" "0 3"; do
set -- $cfg
TSC_GEMM_2CTA=$1 TSC_GEMM_STAGES=$2 timeout 300 python tools/bench_configs.py c3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('2cta=$1 stages=$2', 'hot_ms=%.2f'%d['hot_kernel_ms'], 'TF=%.0f'%d.get('tflops',0), 'frac_burst=%.3f'%d.get('tensor_frac_of_measured_burst',0), 'frac_sust=%.3f'%d.get('tensor_frac_of_measured_sustained',0), 'total_ms=%.2f'%d['device_ms_per_search'])
" | tee -a gpurun_out/gemm10.log
done
echo "== ncu gemm" | tee -a gpurun_out/gemm10.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 2 -c 1 \
  -f -o gpurun_out/prof_gemm python tools/bench_configs.py c3 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/gemm10.log
