#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_diagnostics.py tests/test_gpu_host_program.py -m gpu -x -q > gpurun_out/pytest_diag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_diag.log
tail -15 gpurun_out/pytest_diag.log
for cfg in "512 500 50" "512 1000 250" "256 1000 50" "256 2000 200"; do set -- $cfg
timeout 900 python tools/dump_overlap.py --n $1 --steps $2 --rate $3 > gpurun_out/dump_overlap_$1_$3.json 2> gpurun_out/dump_overlap.err; echo "overlap rc=$?"; cat gpurun_out/dump_overlap_$1_$3.json; tail -3 gpurun_out/dump_overlap.err
done
