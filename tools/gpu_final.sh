#!/bin/bash
# round-end style run on one GPU: build check, smoke, all GPU tests, reference arm, bench, host program timings
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -4 gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-250 gpurun_out/bench_reference.json
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cat gpurun_out/bench_final.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_on_fused_minimum'], d['cpu_baseline']['value'], d['clocks'])"
# the C host program on the reference's own params.txt (the reference needs 6.0 s for this run)
W=$(mktemp -d); mkdir $W/r; printf '0.05\n0.05\n0.05\n0.001\n0.0000000000006\n0.00000000012\n2\n0' > $W/params.txt
( cd $W; /usr/bin/time -f "stock params.txt (50^3, 200 steps, 101 dumps): %e s wall" $GRAFT_REPO_ROOT/fdtd-maxwell-microwave-oven_b200/microwave params.txt > out.txt 2> err.txt; tail -1 err.txt; ls r | wc -l )
printf '0.256\n0.256\n0.256\n0.001\n0.0000000000006\n0.0000000006\n250\n1' > $W/p256.txt
( cd $W; rm -f r/*; FDTD_B200_REPORT=1 /usr/bin/time -f "256^3 x 1000 steps, dump every 250: %e s wall" $GRAFT_REPO_ROOT/fdtd-maxwell-microwave-oven_b200/microwave p256.txt > out.txt 2> err.txt; tail -2 err.txt; du -sh r | cut -f1 )
