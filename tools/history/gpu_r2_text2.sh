#!/bin/bash
# round 2, second text-WHERE call (1 GPU, ~2.5 min): WHERE kernels after the operator decode moved
# to the host and full steps got their own path — GPU tests of both WHERE files, c5w / c5t timings,
# ncu metrics of the three kernels, memcheck over the numeric + text parity tests, then (time
# permitting) the whole GPU suite.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2u}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
echo "== WHERE gpu tests" | tee -a $L
timeout 120 python -m pytest tests/test_where_text.py tests/test_where.py -m gpu -q --timeout 100 2>&1 | tail -12 | tee -a $L
echo "== configs c5w c5t" | tee -a $L
timeout 100 python tools/bench_configs.py c5w c5t 2>&1 | tee gpurun_out/${T}_configs.jsonl | cut -c1-700 | tee -a $L
echo "== ncu: WHERE kernels (c5w: numeric, c5t: text)" | tee -a $L
timeout 100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:"dict_match|where_eval" -c 30 --csv --log-file gpurun_out/${T}_where_ncu.csv \
  python tools/bench_configs.py c5w c5t > gpurun_out/${T}_ncu.log 2>&1
grep -c "dict_match\|where_eval" gpurun_out/${T}_where_ncu.csv | tee -a $L
echo "== memcheck: WHERE kernels (numeric + text), last steps of short columns included" | tee -a $L
timeout 70 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_where_text.py::test_gpu_text_where_equals_oracle tests/test_where.py::test_gpu_filter_where_equals_oracle -m gpu -q -x --timeout 60 -p no:cacheprovider > gpurun_out/${T}_memcheck.txt 2>&1
echo "exit $?" | tee -a $L
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/${T}_memcheck.txt | head -12 | tee -a $L
echo "== all gpu tests" | tee -a $L
timeout 90 python -m pytest tests -m gpu -q --timeout 80 2>&1 | tail -6 | tee -a $L
