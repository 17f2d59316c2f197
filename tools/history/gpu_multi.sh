#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
N=${1:-2}
nvidia-smi -L | tee gpurun_out/multi.log
echo "== pytest multi" | tee -a gpurun_out/multi.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee -a gpurun_out/multi.log
for g in 1 $N; do
echo "== bench gpus=$g" | tee -a gpurun_out/multi.log
if [ "$g" = "1" ]; then
  timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_g$g.err | tee gpurun_out/bench_g$g.json | tee -a gpurun_out/multi.log
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $g --steps 60 --warmup 5 2>gpurun_out/bench_g$g.err | tee gpurun_out/bench_g$g.json | tee -a gpurun_out/multi.log
fi
tail -3 gpurun_out/bench_g$g.err | tee -a gpurun_out/multi.log
done
