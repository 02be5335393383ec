#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --nz 256 --no-e2e --no-cpu --opt kernel=2 --opt strip=1 --opt kchunk=32 --opt prefetch=3 --opt warps_x=1 --opt warps_y=4"
$CMD > gpurun_out/plain_fused.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 3 -c 1 -o gpurun_out/prof_r01_fused $CMD > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fused.log; cat gpurun_out/plain_fused.log | cut -c1-300
