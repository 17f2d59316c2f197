#!/bin/bash
# Round-2 multi-GPU call. Charged N x wall time: keep it SHORT (no CPU-side checks inside
# process groups, small --timeout). Usage: gpurun --gpus 2 --timeout 420 -- 'bash tools/gpu_r2_multi.sh 2'
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
L=gpurun_out/r2_multi_$N.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | tee $L
echo "== sharded parity (nccl + experimental p2p)" | tee -a $L
TSC_TEST_P2P=1 timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5 | tee -a $L
for ex in nccl p2p; do
  echo "== bench gpus=$N exchange=$ex" | tee -a $L
  timeout 240 $TR --nproc-per-node $N --master-port 2956$N bench.py --gpus $N --steps 200 --warmup 5 --exchange $ex 2>gpurun_out/r2_bench_${ex}_$N.err | tee gpurun_out/r2_bench_${ex}_$N.json | tee -a $L
  tail -2 gpurun_out/r2_bench_${ex}_$N.err | tee -a $L
done
if [ "$N" = "8" ]; then
  echo "== c4 (8 shards of 12.5M x 1536 fp16, k=100)" | tee -a $L
  timeout 300 $TR --nproc-per-node 8 --master-port 29571 tools/bench_configs_multi.py c4 2>&1 | tail -2 | tee -a $L
fi
