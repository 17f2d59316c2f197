#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
echo "== pytest gpu" | tee gpurun_out/second.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/second.log
tail -15 gpurun_out/pytest_gpu.log | tee -a gpurun_out/second.log
echo "== bench" | tee -a gpurun_out/second.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/second.log
cat gpurun_out/bench.json | tee -a gpurun_out/second.log
tail -5 gpurun_out/bench.err | tee -a gpurun_out/second.log
echo "== ncu launches" | tee -a gpurun_out/second.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
grep -E "select|scan" gpurun_out/launches.csv | head -6 | tee -a gpurun_out/second.log
echo "== sweep" | tee -a gpurun_out/second.log
timeout 1500 python tools/scan_sweep.py 2>&1 | tee gpurun_out/sweep.log | tail -100 | tee -a gpurun_out/second.log
