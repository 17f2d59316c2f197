#!/bin/bash
# round 2, final 1-GPU evidence call: all GPU tests, smoke, full bench line, reference arm, launch list of the bench
# command, ncu --set full of the dense scan (headline) and of the sparse scan (16 warps)
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2z}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
echo "== gpu tests" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee -a $L
echo "== smoke" | tee -a $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $L
echo "== bench (full line)" | tee -a $L
timeout 900 python bench.py 2>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | cut -c1-400 | tee -a $L
tail -3 gpurun_out/${T}_bench.err | tee -a $L
echo "== bench --no-pipeline (same box)" | tee -a $L
timeout 900 python bench.py --no-pipeline --no-configs --no-cpu-baseline --recall-queries 1 2>>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench_nopipe.json | cut -c1-330 | tee -a $L
echo "== reference arm" | tee -a $L
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench_ref.json | cut -c1-330 | tee -a $L
echo "== launch list of the bench command" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs --recall-queries 0 > gpurun_out/${T}_ncu_bench.log 2>&1
grep -c scan_topk gpurun_out/${T}_launches.csv | tee -a $L
echo "== ncu full: dense scan (headline)" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_kernel -s 6 -c 1 -o gpurun_out/${T}_scan_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --recall-queries 0 > gpurun_out/${T}_ncu_scan.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_scan_full.ncu-rep > gpurun_out/${T}_scan_ncu_summary.txt 2>&1
rm -f gpurun_out/${T}_scan_full.ncu-rep
head -12 gpurun_out/${T}_scan_ncu_summary.txt | tee -a $L
echo "== ncu full: sparse scan (C5, 16 warps)" | tee -a $L
timeout 600 ncu --set full --clock-control none -k regex:scan_topk_sparse -s 4 -c 1 -o gpurun_out/${T}_sparse_full python tools/bench_configs.py c5 > gpurun_out/${T}_ncu_sparse.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_sparse_full.ncu-rep > gpurun_out/${T}_sparse_ncu_summary.txt 2>&1
rm -f gpurun_out/${T}_sparse_full.ncu-rep
head -14 gpurun_out/${T}_sparse_ncu_summary.txt | tee -a $L
