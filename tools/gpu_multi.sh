#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -25 gpurun_out/pytest_multi.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 3 --no-e2e > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
  echo "bench n=$n rc=$?"; cat gpurun_out/bench_n$n.json | cut -c1-400; tail -3 gpurun_out/bench_n$n.err
done
