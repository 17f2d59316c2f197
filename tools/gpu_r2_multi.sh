#!/bin/bash
# Round-2 multi-GPU call. Charged N x wall time: keep it SHORT (no CPU-side checks inside
# process groups, small --timeout). Usage: gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_r2_multi.sh 2'
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-r2m}
L=gpurun_out/${TAG}_$N.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | tee $L
nvidia-smi topo -m 2>&1 | head -12 | tee -a $L
echo "== sharded parity: process per GPU (p2p push + nccl) and single-process group handle" | tee -a $L
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -m gpu -q --timeout 300 2>&1 | tail -8 | tee -a $L
EXS="p2p nccl"; [ "$N" = "8" ] && EXS="p2p"
for ex in $EXS; do
  echo "== bench gpus=$N exchange=$ex" | tee -a $L
  timeout 300 $TR --nproc-per-node $N --master-port 2956$N bench.py --gpus $N --steps 200 --warmup 5 --exchange $ex --recall-queries 2 2>gpurun_out/${TAG}_bench_${ex}_$N.err | tee gpurun_out/${TAG}_bench_${ex}_$N.json | cut -c1-3000 | tee -a $L
  tail -2 gpurun_out/${TAG}_bench_${ex}_$N.err | tee -a $L
done
echo "== single-process group handle, host buffers (tools/bench_group.py)" | tee -a $L
timeout 300 python tools/bench_group.py $N 10000000 200 2>&1 | tail -2 | tee gpurun_out/${TAG}_group_$N.json | tee -a $L
if [ "$N" = "8" ]; then
  echo "== c4 (8 shards of 12.5M x 1536 fp16, k=100)" | tee -a $L
  timeout 420 $TR --nproc-per-node 8 --master-port 29571 tools/bench_configs_multi.py c4 --exchange p2p --check 2>&1 | tail -2 | tee gpurun_out/${TAG}_c4_$N.json | tee -a $L
  echo "== c5 (4 shards of 12.5M x 384 fp32, 10% mask, k=10) on GPUs 0-3" | tee -a $L
  timeout 300 $TR --nproc-per-node 4 --master-port 29572 tools/bench_configs_multi.py c5 --exchange p2p --check 2>&1 | tail -2 | tee gpurun_out/${TAG}_c5_4.json | tee -a $L
fi
