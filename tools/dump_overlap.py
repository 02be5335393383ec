#!/usr/bin/env python
"""BASELINE.json configs[4]: a production-length run with dumps every 50 steps -- does output stall the
stepping loop?  Times fdtd_propagate with a sink that consumes every variable (memcpy-speed sink)
against the same run without dumps.      python tools/dump_overlap.py [--n 512] [--steps 1000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_b200 as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--rate", type=int, default=50)
args = ap.parse_args()

dt = 6e-13
sim = np.float32(dt * (args.steps - 0.5))
p = F.make_params(args.n * 1e-3, args.n * 1e-3, args.n * 1e-3, 1e-3, dt, float(sim), args.rate, 1)
assert p.dims() == (args.n,) * 3
steps = F.step_count(p)
cells = args.n ** 3
out = {"grid": args.n, "steps": steps, "sampling_rate": args.rate}
for dumps in (False, True):
    with F.Context(p) as ctx:
        ctx.run(3, 0.0)           # warm-up (allocates the second state copy)
        ctx.sync()
        seen = {"vars": 0, "bytes": 0, "files": 0}

        def on_variable(name, arr):
            seen["vars"] += 1
            seen["bytes"] += arr.nbytes

        t0 = time.perf_counter()
        n, _ = ctx.propagate(on_begin=lambda it, dims, k0: seen.__setitem__("files", seen["files"] + 1),
                             on_variable=on_variable, dumps=dumps)
        dt_s = time.perf_counter() - t0
        key = "with_dumps" if dumps else "no_dumps"
        out[key] = {"seconds": dt_s, "gcell_s": cells * n / dt_s / 1e9, "steps": n, **seen}
out["slowdown"] = out["with_dumps"]["seconds"] / out["no_dumps"]["seconds"]
print(json.dumps(out))
