#!/bin/bash
# round-end style run on one GPU: smoke, all GPU tests, reference arm, default bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -4 gpurun_out/pytest_gpu_final.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cat gpurun_out/bench_final.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_on_fused_minimum'], d['cpu_baseline']['value'], d['clocks'])"
python - <<'PY'
import os, subprocess, tempfile, time
exe = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "fdtd-maxwell-microwave-oven_b200", "microwave")
w = tempfile.mkdtemp(); os.mkdir(w + "/r")
open(w + "/params.txt", "w").write("0.05\n0.05\n0.05\n0.001\n0.0000000000006\n0.00000000012\n2\n0")
out = []
for rep in ("first run", "second run"):
    t0 = time.time(); r = subprocess.run([exe, "params.txt"], cwd=w, capture_output=True, text=True); dt = time.time() - t0
    out.append(f"stock params.txt (50^3, 200 steps, validation mode, 101 dumps), {rep}: rc={r.returncode} {dt:.2f} s wall, {len([f for f in os.listdir(w + '/r') if f.endswith('.raw')])} dump files (reference on one CPU core: 6.0 s, SURVEY.md 6)")
open(w + "/p256.txt", "w").write("0.256\n0.256\n0.256\n0.001\n0.0000000000006\n0.0000000006\n250\n1")
for f in os.listdir(w + "/r"): os.remove(w + "/r/" + f)
t0 = time.time(); r = subprocess.run([exe, "p256.txt"], cwd=w, capture_output=True, text=True, env=dict(os.environ, FDTD_B200_REPORT="1")); dt = time.time() - t0
out.append(f"256^3 x 1000 steps, computation mode, dump every 250 (5 dumps written to disk): rc={r.returncode} {dt:.2f} s wall; {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}")
open("gpurun_out/host_program_timings.txt", "w").write("\n".join(out) + "\n"); print("\n".join(out))
PY
