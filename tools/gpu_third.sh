#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err; echo "bench rc=$?"
cat gpurun_out/bench_tma.json; tail -3 gpurun_out/bench_tma.err
CMD="python bench.py --steps 2 --warmup 3 --nz 256 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_tma.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_fused_tma -s 3 -c 1 -o gpurun_out/prof_r01_tma $CMD > gpurun_out/ncu_tma.log 2>&1
echo "ncu rc=$?"
$CMD > gpurun_out/plain_tma.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r01_tma.csv $CMD > gpurun_out/ncu_tma1.log 2>&1
echo "ncu launches rc=$?"
