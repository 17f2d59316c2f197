#!/bin/bash
# round 2, text-WHERE evidence call (1 GPU, ~3 min): the new GPU tests first, then the whole GPU
# suite, the c5w / c5t configs, an ncu pass over dict_match_kernel + where_eval_kernel<true>,
# and a short headline bench as a sanity check that the scan path is where it was.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2t}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
echo "== text WHERE gpu tests" | tee -a $L
timeout 240 python -m pytest tests/test_where_text.py tests/test_where.py -m gpu -q --timeout 120 2>&1 | tail -15 | tee -a $L
echo "== all gpu tests" | tee -a $L
timeout 400 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8 | tee -a $L
echo "== configs c5w c5t" | tee -a $L
timeout 200 python tools/bench_configs.py c5w c5t 2>&1 | tee gpurun_out/${T}_configs.jsonl | cut -c1-600 | tee -a $L
echo "== ncu: text WHERE kernels" | tee -a $L
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:"dict_match|where_eval" -c 12 --csv --log-file gpurun_out/${T}_where_ncu.csv \
  python tools/bench_configs.py c5t > gpurun_out/${T}_ncu.log 2>&1
grep -c "dict_match\|where_eval" gpurun_out/${T}_where_ncu.csv | tee -a $L
echo "== C++ demo (GPU half)" | tee -a $L
g++ -std=c++17 -I include examples/vector_store_demo.cc -L tostore_b200 -ltostore_cuda -Wl,-rpath,$PWD/tostore_b200 -o /tmp/vsdemo 2>&1 | tail -3 | tee -a $L
timeout 60 /tmp/vsdemo 2>&1 | tail -22 | tee -a $L
echo "== short headline bench" | tee -a $L
timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-configs --recall-queries 1 2>gpurun_out/${T}_bench.err | tee gpurun_out/${T}_bench.json | cut -c1-420 | tee -a $L
tail -2 gpurun_out/${T}_bench.err | tee -a $L
echo "== memcheck: WHERE kernels (numeric + text), last steps of short columns included" | tee -a $L
timeout 100 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_where_text.py::test_gpu_text_where_equals_oracle tests/test_where.py::test_gpu_filter_where_equals_oracle -m gpu -q -x --timeout 90 -p no:cacheprovider > gpurun_out/${T}_memcheck.txt 2>&1
echo "exit $?" | tee -a $L
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/${T}_memcheck.txt | head -12 | tee -a $L
