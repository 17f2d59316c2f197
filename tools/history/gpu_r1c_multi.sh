#!/bin/bash
# multi-GPU call: sharded parity test, headline bench at 1/2/4/../N GPUs, sharded configs C4/C5
set +e
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-2}
L=gpurun_out/r1c_multi_$N.log
nvidia-smi -L | tee $L
nvidia-smi topo -m 2>/dev/null | head -12 | tee -a $L
echo "== pytest multi" | tee -a $L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 | tee -a $L
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for g in 1 2 4 8; do
  [ $g -gt $N ] && break
  echo "== bench gpus=$g" | tee -a $L
  if [ "$g" = "1" ]; then
    timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/r1c_bench_g$g.err | tee gpurun_out/r1c_bench_g$g.json | tee -a $L
  else
    timeout 600 $TR --nproc-per-node $g --master-port 2953$g bench.py --gpus $g --steps 100 --warmup 5 2>gpurun_out/r1c_bench_g$g.err | tee gpurun_out/r1c_bench_g$g.json | tee -a $L
  fi
  tail -2 gpurun_out/r1c_bench_g$g.err | tee -a $L
done
echo "== sharded configs" | tee -a $L
G4=$(( N < 4 ? N : 4 ))
timeout 900 $TR --nproc-per-node $N --master-port 29541 tools/bench_configs_multi.py c4 --check 2>gpurun_out/r1c_c4.err | tee -a gpurun_out/r1c_multi_configs_$N.jsonl | tee -a $L
tail -2 gpurun_out/r1c_c4.err | tee -a $L
timeout 900 $TR --nproc-per-node $G4 --master-port 29542 tools/bench_configs_multi.py c5 --check 2>gpurun_out/r1c_c5.err | tee -a gpurun_out/r1c_multi_configs_$N.jsonl | tee -a $L
tail -2 gpurun_out/r1c_c5.err | tee -a $L
echo "== NCCL algo for the all-gather" | tee -a $L
NCCL_DEBUG=INFO timeout 300 $TR --nproc-per-node $N --master-port 29543 tools/bench_configs_multi.py c2 --steps 5 2>&1 | grep -E "NVLS|Channel|via P2P|Connected all|\"config\"" | head -12 | tee -a $L
