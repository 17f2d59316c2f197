#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
echo "== c3 epilogue experiments (EXP): 16 ldtm-only, 32 compute-only, 64 no-norm-loads" | tee gpurun_out/gemm8.log
for cfg in 0 2 16 32 64 80 96; do
TSC_GEMM_EXP=$cfg TSC_GEMM_L2PF=0 timeout 300 python tools/bench_configs.py c3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print('exp=$cfg', 'hot_ms=%.2f'%d['hot_kernel_ms'], 'TF=%.0f'%d.get('tflops',0), 'frac_burst=%.3f'%d.get('tensor_frac_of_measured_burst',0))
" | tee -a gpurun_out/gemm8.log
done
