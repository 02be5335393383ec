export NCU_TAG=k4_default NCU_KERNEL=k_step2 NCU_SKIP=2
tools/gpu.sh ncu --steps 4 --warmup 4 --nz 256 --no-e2e --no-cpu --no-selfcheck
python bench.py --workload dumps50 > gpurun_out/dumps50.json 2> gpurun_out/dumps50.err; echo "dumps50 rc=$?"; cat gpurun_out/dumps50.json
python - <<'PY'
import os, subprocess, tempfile, time
exe = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "fdtd-maxwell-microwave-oven_b200", "microwave")
w = tempfile.mkdtemp(); os.mkdir(w + "/r")
open(w + "/params.txt", "w").write("0.05\n0.05\n0.05\n0.001\n0.0000000000006\n0.00000000012\n2\n0")
out = []
for rep in ("first run", "second run"):
    t0 = time.time(); r = subprocess.run([exe, "params.txt"], cwd=w, capture_output=True, text=True, env=dict(os.environ, FDTD_B200_REPORT="1")); dt = time.time() - t0
    out.append(f"stock params.txt (50^3, 200 steps, validation mode, 101 Silo dumps), {rep}: rc={r.returncode} {dt:.2f} s wall, {len([f for f in os.listdir(w + '/r') if f.endswith('.silo')])} .silo files; {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''} (reference on one CPU core: 6.0 s, SURVEY.md 6)")
r = subprocess.run([exe, "params.txt"], cwd=w, capture_output=True, text=True, env=dict(os.environ, FDTD_B200_REPORT="1", FDTD_B200_NO_DUMPS="1"))
out.append(f"same without dumps: {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}")
open(w + "/p256.txt", "w").write("0.256\n0.256\n0.256\n0.001\n0.0000000000006\n0.0000000006\n250\n1")
for f in os.listdir(w + "/r"): os.remove(w + "/r/" + f)
t0 = time.time(); r = subprocess.run([exe, "p256.txt"], cwd=w, capture_output=True, text=True, env=dict(os.environ, FDTD_B200_REPORT="1")); dt = time.time() - t0
out.append(f"256^3 x 1000 steps, computation mode, Silo dump every 250 (5 files of 805 MB): rc={r.returncode} {dt:.2f} s wall; {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}")
open("gpurun_out/host_program_timings.txt", "w").write("\n".join(out) + "\n"); print("\n".join(out))
PY
