#!/bin/bash
# round 2, call O (1 GPU): column growth (VMM row block), tensor stats, full test suite, bench sanity
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2o}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
timeout 900 python -m pytest tests/test_gpu_growth.py -m gpu -q --timeout 300 -x 2>&1 | tail -15 | tee -a $L
echo "== all gpu tests" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee -a $L
echo "== bench (full line)" | tee -a $L
timeout 900 python bench.py 2>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | cut -c1-600 | tee -a $L
tail -3 gpurun_out/${T}_bench.err | tee -a $L
echo "== reference arm" | tee -a $L
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench_ref.json | cut -c1-600 | tee -a $L
echo "== smoke" | tee -a $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee -a $L
