#!/bin/bash
# 8-GPU session on ONE box (run under gpurun --gpus 8): same-box scaling table, weak and strong
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|^cpu(s)\|model name"; free -g | head -2; } > gpurun_out/box8_topology.txt 2>&1
if [ "$1" == "tests" ]; then
  FDTD_MULTI_QUICK=1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "8-peer or 8-nccl" > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi_8.log
fi
export BENCH_TAG=scale_n1; tools/gpu.sh bench --steps 20 --warmup 5 --no-e2e --no-cpu
for n in 2 4; do export BENCH_TAG=scale_n$n; tools/gpu.sh benchn $n --steps 20 --warmup 5 --no-e2e; done
export BENCH_TAG=scale_n8; tools/gpu.sh benchn 8 --steps 20 --warmup 5
export BENCH_TAG=strong2048_n4; tools/gpu.sh benchn 4 --steps 40 --workload strong2048 --no-e2e
export BENCH_TAG=strong2048_n8; tools/gpu.sh benchn 8 --steps 40 --workload strong2048 --no-e2e
