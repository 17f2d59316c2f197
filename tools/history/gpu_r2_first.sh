#!/bin/bash
# First GPU call of the next round: everything written after round 1 ran out of GPU budget.
# 1 GPU, a few minutes. Keep --timeout small; never wrap multi-minute CPU work in a
# multi-GPU call (round 1 lost 130 GPU-minutes that way).
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_first.log
nvidia-smi -L | tee $L
which dart flutter 2>&1 | tee -a $L      # a Dart SDK on the box would let the real reference run
echo "== pytest gpu (default)" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee -a $L
echo "== pytest gpu (experimental paths)" | tee -a $L
TSC_TEST_TF32=1 TSC_TEST_SPARSE_PF=1 TSC_TEST_PROPERTY=1 timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_round1_late_gpu.py -m gpu -q -k "tf32 or sparse or property" 2>&1 | tail -12 | tee -a $L
echo "== C example" | tee -a $L
gcc -std=c99 -I include examples/minimal.c -L tostore_b200 -ltostore_cuda -Wl,-rpath,$PWD/tostore_b200 -lm -o /tmp/minimal && /tmp/minimal 2>&1 | tee -a $L
echo "== C++ host layer demo" | tee -a $L
g++ -std=c++17 -I include examples/vector_store_demo.cc -L tostore_b200 -ltostore_cuda -Wl,-rpath,$PWD/tostore_b200 -o /tmp/vsdemo && /tmp/vsdemo 2>&1 | tail -16 | tee -a $L
echo "== bench" | tee -a $L
timeout 600 python bench.py 2>gpurun_out/r2_bench.err | tee gpurun_out/r2_bench.json | tee -a $L
echo "== configs (new paths)" | tee -a $L
timeout 900 python tools/bench_configs.py c5 c5pf c5w c2t 2>&1 | tee gpurun_out/r2_configs.jsonl | tee -a $L
echo "== native loader throughput (1M x 768 fp32 = 3.3 GB on disk)" | tee -a $L
timeout 600 python tools/bench_loader.py --rows 1000000 --dims 768 2>&1 | tail -4 | tee -a $L
echo "== ncu where kernel" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:where_eval -c 3 python tools/bench_configs.py c5w 2>&1 | grep -E "where_eval|gpu__time|dram__" | tee -a $L
echo "== ncu full: sparse scan (is C5 bound by the bitmap stall or by DRAM page misses?)" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_sparse -s 3 -c 1 -o gpurun_out/r2_sparse_full python tools/bench_configs.py c5 > gpurun_out/r2_ncu_sparse.log 2>&1
tail -2 gpurun_out/r2_ncu_sparse.log | tee -a $L
TSC_SCAN_SPARSE_PF=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_topk_sparse -s 3 -c 1 -o gpurun_out/r2_sparse_pf_full python tools/bench_configs.py c5 > gpurun_out/r2_ncu_sparse_pf.log 2>&1
tail -2 gpurun_out/r2_ncu_sparse_pf.log | tee -a $L
echo "== microbench: HBM bandwidth for sparse 1.5 KB row reads (roofline denominator for C5)" | tee -a $L
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench/gather_bw.cu -o /tmp/gather_bw && timeout 300 /tmp/gather_bw 12500000 1536 2>&1 | tee -a $L
