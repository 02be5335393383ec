#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep.py --nz 512 --steps 5 --only 2 > gpurun_out/sweep_fused.txt 2> gpurun_out/sweep_fused.err; echo "sweep rc=$?"; tail -3 gpurun_out/sweep_fused.err
