#!/bin/bash
# round 2, call Q (1 GPU): all GPU tests after the packed result block; e2e of the small configs
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2q}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee -a $L
timeout 600 python tools/bench_configs.py c1 c5 2>&1 | cut -c1-420 | tee -a $L
timeout 600 python bench.py --rows 1250000 --steps 400 --no-configs --no-cpu-baseline --recall-queries 1 2>>gpurun_out/${T}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'ms_per_step', round(d['ms_per_step'],4), 'e2e', d['e2e'], d['recall_check']['bit_exact'])" | tee -a $L
