#!/bin/bash
# round-1 session c, call 1: verify restored state (gpu tests, bench, configs, ncu launch list)
set +e
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r1c_1.log
nvidia-smi -L | tee $L
echo "== pytest gpu" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee -a $L
echo "== smoke" | tee -a $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $L
echo "== bench" | tee -a $L
timeout 600 python bench.py 2>gpurun_out/r1c_bench.err | tee gpurun_out/r1c_bench.json | tee -a $L
tail -3 gpurun_out/r1c_bench.err | tee -a $L
echo "== bench reference" | tee -a $L
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/r1c_bench_ref.json | tee -a $L
echo "== configs" | tee -a $L
timeout 900 python tools/bench_configs.py c1 c2 c2b c3 c3s c4 c5 2>&1 | tee gpurun_out/r1c_configs.jsonl | tee -a $L
echo "== ncu launch list (bench)" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_bench.log 2>&1
tail -2 gpurun_out/r1c_ncu_bench.log | tee -a $L
echo "== ncu full scan kernel" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/r1c_scan_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_scan.log 2>&1
tail -2 gpurun_out/r1c_ncu_scan.log | tee -a $L
nproc | tee -a $L
