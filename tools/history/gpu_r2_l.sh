#!/bin/bash
# round 2, call L (2 GPUs): all GPU tests (pipelining without range launch, group worker threads), N=2 bench, group bench
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
T=r2l
L=gpurun_out/${T}_$N.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | tee $L
echo "== all gpu tests" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -12 | tee -a $L
echo "== bench gpus=$N (p2p, pipelined device loop)" | tee -a $L
timeout 300 $TR --nproc-per-node $N --master-port 2956$N bench.py --gpus $N --steps 200 --warmup 5 --recall-queries 2 2>gpurun_out/${T}_bench_$N.err | tee gpurun_out/${T}_bench_$N.json | cut -c1-3200 | tee -a $L
tail -2 gpurun_out/${T}_bench_$N.err | tee -a $L
echo "== single-process group handle, host buffers (tools/bench_group.py)" | tee -a $L
timeout 300 python tools/bench_group.py $N 10000000 200 2>&1 | tail -2 | tee gpurun_out/${T}_group_$N.json | tee -a $L
echo "== 1 GPU bench (pipelined)" | tee -a $L
timeout 600 python bench.py --no-configs --no-cpu-baseline --recall-queries 1 2>>gpurun_out/${T}_bench_$N.err | cut -c1-330 | tee -a $L
