#!/bin/bash
# First GPU pass: smoke, parity tests, bench, config sweep, ncu captures.
# Everything is logged under gpurun_out/ (merged back by gpurun).
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" | tee gpurun_out/first.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/first.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/first.log
echo "== pytest gpu" | tee -a gpurun_out/first.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/first.log
tail -15 gpurun_out/pytest_gpu.log | tee -a gpurun_out/first.log
echo "== bench" | tee -a gpurun_out/first.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" | tee -a gpurun_out/first.log
cat gpurun_out/bench.json | tee -a gpurun_out/first.log
tail -5 gpurun_out/bench.err | tee -a gpurun_out/first.log
echo "== sweep" | tee -a gpurun_out/first.log
for cfg in "8 0 0 8192" "8 0 1 8192" "8 0 4 16384" "16 0 1 8192" "16 0 2 8192" "4 0 2 8192" "4 0 4 16384" "8 3 2 8192" "8 2 2 8192" "16 4 1 8192"; do
  set -- $cfg
  TSC_SCAN_WARPS=$1 TSC_SCAN_STAGES=$2 TSC_SCAN_ROWS=$3 TSC_SCAN_STAGE_BYTES=$4 \
    timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | \
    python -c "import sys,json; l=json.loads(sys.stdin.read()); r=l['roofline']; print('cfg $cfg', 'qps=%.1f'%l['value'], 'scan_ms=%.3f'%r['kernel_ms'], 'GB/s=%.0f'%r['achieved'], 'frac=%.3f'%r['frac'], 'e2e=%.1f'%l['e2e']['value'])" | tee -a gpurun_out/first.log
done
echo "== ncu launches" | tee -a gpurun_out/first.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/first.log
echo "== ncu full" | tee -a gpurun_out/first.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 2 \
  -f -o gpurun_out/prof_scan python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a gpurun_out/first.log
ls -la gpurun_out | tee -a gpurun_out/first.log
