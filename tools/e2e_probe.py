"""End-to-end timing probe on one GPU: plain upload / download times, then fdtd_run_hosted for a few
chunk sizes (wall clock, pinned host arrays).  python tools/e2e_probe.py [--n 1024] [--steps 20]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_b200 as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--chunks", default="4,8,16,32")
args = ap.parse_args()
n = args.n
p = F.make_params(n * 1e-3, n * 1e-3, n * 1e-3, 1e-3, 6e-13, 1e-9, 1 << 30, 1)
cells = n ** 3
with F.Context(p) as ctx:
    host = F.PinnedArrays(p)
    nbytes = sum(a.nbytes for a in host.arrays.values())
    ctx.fill_test_pattern(7)
    ctx.download_slab(host.arrays)
    ctx.run(2, 0.0)
    ctx.sync()
    for what, fn in (("upload", lambda: ctx.upload_slab(host.arrays)), ("download", lambda: ctx.download_slab(host.arrays))):
        t0 = time.perf_counter(); fn(); ctx.sync(); dt = time.perf_counter() - t0
        print(json.dumps({"what": what, "seconds": dt, "GBps": nbytes / dt / 1e9}), flush=True)
    t0 = time.perf_counter(); ctx.run(args.steps, 0.0); ctx.sync(); dt = time.perf_counter() - t0
    print(json.dumps({"what": f"run({args.steps}) device-resident", "seconds": dt}), flush=True)
    for pipeline, chunk in [(0, 0)] + [(1, int(c)) for c in args.chunks.split(",")]:
        ctx.set_option("host_pipeline", pipeline)
        ctx.set_option("host_chunk", chunk)
        ctx.fill_test_pattern(7)
        ctx.download_slab(host.arrays)
        ctx.sync()
        t0 = time.perf_counter(); ctx.run_hosted(host.arrays, args.steps, 0.0); dt = time.perf_counter() - t0
        print(json.dumps({"what": "run_hosted", "pipeline": pipeline, "host_chunk": chunk, "seconds": dt,
                          "gcell_s": cells * args.steps / dt / 1e9}), flush=True)
    host.close()
