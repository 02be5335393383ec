#!/bin/bash
# BASELINE.json configs[3]: 2048^3 on 8 GPUs, and strong scaling of a fixed 2048 x 2048 x 1024 cavity on 2/4/8
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
run() { # name n extra-args
  local name=$1 n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 40 --warmup 4 --no-e2e --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"; python -c "
import json
for l in open('gpurun_out/$name.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d['config']['workload'][:24], d['hbm_bytes_per_gpu'])
"; grep -iE "error|Traceback" gpurun_out/$name.err | head -3
}
[ $NG -ge 8 ] && run cube2048_n8 8 --nxy 2048 --nz 256
for n in 2 4 8; do [ $n -le $NG ] && run strong2048x1024_n$n $n --nxy 2048 --nz-total 1024; done
