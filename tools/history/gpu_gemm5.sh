#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
echo "== experiments (TS kernel, results invalid for exp != 0)" | tee gpurun_out/gemm5.log
for cfg in "0 8" "1 8" "2 8" "3 8" "4 8" "6 8"; do
set -- $cfg
TSC_GEMM_TS=1 TSC_GEMM_EXP=$1 TSC_GEMM_STAGES=$2 timeout 300 python tools/bench_configs.py c3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('exp=$1 stages=$2', 'hot_ms=%.2f'%d['hot_kernel_ms'], 'TF=%.0f'%d.get('tflops',0), 'frac_burst=%.3f'%d.get('tensor_frac_of_measured_burst',0))
" | tee -a gpurun_out/gemm5.log
done
