#!/bin/bash
# round 2, last call (1 GPU, ~75 s): shipping library with the L2 prefetch at distance 1 and the
# open-addressing dictionary — WHERE GPU tests; then the prefetch-distance A/B and the append
# rates through the diagnostics build; then as much of the whole GPU suite as the budget allows.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2v}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
echo "== WHERE gpu tests (shipping library)" | tee -a $L
timeout 60 python -m pytest tests/test_where_text.py tests/test_where.py -m gpu -q --timeout 50 2>&1 | tail -8 | tee -a $L
echo "== prefetch A/B + append rates (diagnostics build)" | tee -a $L
timeout 60 python tools/where_ab.py 2>&1 | tee gpurun_out/${T}_where_ab.json | tee -a $L
echo "== all gpu tests" | tee -a $L
timeout 70 python -m pytest tests -m gpu -q --timeout 60 2>&1 | tail -6 | tee -a $L
