#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "bench rc=$?"
cat gpurun_out/bench_e2e.json; tail -5 gpurun_out/bench_e2e.err
CMD="python bench.py --steps 3 --warmup 3 --nz 256 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_update -s 8 -c 2 -o gpurun_out/prof_r01_split $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu2.log
