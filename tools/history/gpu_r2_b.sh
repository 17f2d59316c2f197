#!/bin/bash
# round 2, call B (N GPUs, default 2): sharded exchange (p2p fused + nccl), the single-process
# group, bench at N. Keep it short: charged N x wall time.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
L=gpurun_out/r2b_$N.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | tee $L
echo "== multi / group tests" | tee -a $L
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -m gpu -q --timeout 300 2>&1 | tail -30 | tee -a $L
echo "== group bench (one process, $N GPUs)" | tee -a $L
timeout 300 python tools/bench_group.py $N 10000000 200 2>&1 | tail -3 | tee gpurun_out/r2b_group_$N.json | tee -a $L
for ex in p2p nccl; do
  echo "== bench gpus=$N exchange=$ex" | tee -a $L
  timeout 300 $TR --nproc-per-node $N --master-port 2956$N bench.py --gpus $N --steps 200 --warmup 5 --exchange $ex --recall-queries 2 2>gpurun_out/r2b_bench_${ex}_$N.err | tee gpurun_out/r2b_bench_${ex}_$N.json | tee -a $L
  tail -3 gpurun_out/r2b_bench_${ex}_$N.err | tee -a $L
done
