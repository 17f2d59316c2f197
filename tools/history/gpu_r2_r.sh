#!/bin/bash
# round 2, call R (1 GPU): small batches (2 / 4 / 8 queries): multi-query scan vs the tensor path
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2r}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
FMT="
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], 'path', d['path'], 'device_ms', round(d['device_ms_per_search'],3), 'hot_ms', round(d['hot_kernel_ms'],3), 'cert', d['certified'], d['retried'], d['uncertified'])
    else: print(l.strip()[:200])"
for cfg in c2q c3q; do
  echo "== $cfg, scan path (default: tensor path from 9 queries)" | tee -a $L
  timeout 600 python tools/bench_configs.py --diag $cfg 2>&1 | python -c "$FMT" | tee -a $L
  echo "== $cfg, tensor path forced (TSC_GEMM_MIN_NQ=2)" | tee -a $L
  TSC_GEMM_MIN_NQ=2 timeout 600 python tools/bench_configs.py --diag $cfg 2>&1 | python -c "$FMT" | tee -a $L
done
