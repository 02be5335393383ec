#!/bin/bash
# 8-GPU validation (run under gpurun --gpus 8): world-8 slab tests, weak and strong scaling benches
mkdir -p gpurun_out
FDTD_MULTI_QUICK=1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "8-peer or 8-nccl" > gpurun_out/pytest_multi_8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi_8.log
export BENCH_TAG=n8_weak; tools/gpu.sh benchn 8 --steps 20 --warmup 5
export BENCH_TAG=n8_weak_nccl; tools/gpu.sh benchn 8 --steps 20 --warmup 5 --transport nccl --no-e2e --no-selfcheck
export BENCH_TAG=n8_strong2048; tools/gpu.sh benchn 8 --steps 40 --workload strong2048 --no-e2e
export BENCH_TAG=n8_cube2048; tools/gpu.sh benchn 8 --steps 20 --workload cube2048 --no-e2e
