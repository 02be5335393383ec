#!/bin/bash
mkdir -p gpurun_out
EXE=$GRAFT_REPO_ROOT/fdtd-maxwell-microwave-oven_b200/microwave
W=$(mktemp -d); mkdir $W/r; printf '0.05\n0.05\n0.05\n0.001\n0.0000000000006\n0.00000000012\n2\n0' > $W/params.txt
cd $W
t0=$(date +%s.%N); $EXE params.txt > out.txt 2> err.txt; rc=$?; t1=$(date +%s.%N)
echo "stock params.txt (50^3, 200 steps, validation mode, 101 dumps): rc=$rc $(echo "$t1 - $t0" | bc) s wall, $(ls r/*.raw | wc -l) dump files" | tee $GRAFT_REPO_ROOT/gpurun_out/host_program_timings.txt
t0=$(date +%s.%N); $EXE params.txt > out.txt 2> err.txt; t1=$(date +%s.%N)
echo "same, second run (driver warm): $(echo "$t1 - $t0" | bc) s wall" | tee -a $GRAFT_REPO_ROOT/gpurun_out/host_program_timings.txt
printf '0.256\n0.256\n0.256\n0.001\n0.0000000000006\n0.0000000006\n250\n1' > p256.txt; rm -f r/*
t0=$(date +%s.%N); FDTD_B200_REPORT=1 $EXE p256.txt > out.txt 2> err.txt; t1=$(date +%s.%N)
echo "256^3 x 1000 steps, computation mode, dump every 250 (5 dumps, $(du -sh r | cut -f1)): $(echo "$t1 - $t0" | bc) s wall; $(tail -1 err.txt)" | tee -a $GRAFT_REPO_ROOT/gpurun_out/host_program_timings.txt
cd $GRAFT_REPO_ROOT
if [ $(nvidia-smi -L | wc -l) -ge 2 ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_2.log; tail -4 gpurun_out/pytest_multi_2.log
fi
