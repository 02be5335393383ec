#!/bin/bash
# first GPU call: environment probe, parity tests, variant sweep, short bench
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
  nproc; free -g | head -2; cat /sys/fs/cgroup/memory.max 2>/dev/null; ulimit -l
} > gpurun_out/env.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep.py --nz 512 --steps 5 > gpurun_out/sweep.txt 2> gpurun_out/sweep.err; echo "sweep rc=$?"
sort -t: -k7 gpurun_out/sweep.txt | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench rc=$?"
cat gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
