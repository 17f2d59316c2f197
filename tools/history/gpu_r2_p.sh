#!/bin/bash
# round 2, call P (1 GPU): pipelined vs unpipelined device loop at 10M / 5M / 2.5M / 1.25M rows on ONE box; PK test
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2p}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
timeout 600 python -m pytest tests/test_round1_late_gpu.py tests/test_gpu_growth.py -m gpu -q --timeout 300 2>&1 | tail -3 | tee -a $L
for rows in 10000000 5000000 2500000 1250000; do
  for mode in "" "--no-pipeline"; do
    echo "== rows=$rows $mode" | tee -a $L
    timeout 600 python bench.py --rows $rows --steps 200 --no-configs --no-cpu-baseline --recall-queries 0 $mode 2>>gpurun_out/${T}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms_per_step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'launches', d['gpu_launches'], 'sm_mhz', d['clocks']['sm_mhz'])" | tee -a $L
  done
done
tail -3 gpurun_out/${T}.err | tee -a $L
