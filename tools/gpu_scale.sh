#!/bin/bash
# scaling run on one box: slab parity tests at every world size, then bench at N = 1, 2, 4, 8
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
FDTD_MULTI_QUICK=1 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_$NG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_$NG.log
tail -5 gpurun_out/pytest_multi_$NG.log
for n in 1 2 4 8; do
  [ $n -gt $NG ] && continue
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "bench n=$n rc=$?"; python -c "
import json,sys
for l in open('gpurun_out/scale_n$n.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e'] and d['e2e'].get('value'), d['clocks'])
"; grep -iE "error|Traceback" gpurun_out/scale_n$n.err | head -3
done
