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
out = {"grid": args.n, "steps": steps, "sampling_rate": args.rate,
       "sink": "counts variables only (the pinned buffer is handed over, nothing is copied again)"}
for dumps in (False, True):
    with F.Context(p) as ctx:
        ctx.run(3, 0.0)           # warm-up (allocates the second state copy)
        ctx.sync()
        seen = {"vars": 0, "bytes": 0, "files": 0}

        nbytes = cells * 8

        def count_variable(user, name, data, count):
            seen["vars"] += 1
            seen["bytes"] += count * 8
            return 0

        import ctypes as C
        sink = F.DumpSink(None, F._BEGIN(lambda u, it, dims, k0: seen.__setitem__("files", seen["files"] + 1) or 0),
                          F._VARIABLE(count_variable), F._END(lambda u: 0))
        st, tc = C.c_size_t(), C.c_double()
        for rep in ("first call (allocates scratch + pinned buffers)", "second call"):
            seen.update(vars=0, bytes=0, files=0)
            t0 = time.perf_counter()
            if dumps:
                F._check(F.lib.fdtd_propagate(ctx._h, C.byref(sink), C.byref(st), C.byref(tc)))
                n = int(st.value)
            else:
                n, _ = ctx.propagate(dumps=False)
            dt_s = time.perf_counter() - t0
            out.setdefault("first_call_seconds", {})["with_dumps" if dumps else "no_dumps"] = out.get("_prev", dt_s) if rep.startswith("second") else dt_s
            out["_prev"] = dt_s
        out.pop("_prev", None)
        key = "with_dumps" if dumps else "no_dumps"
        out[key] = {"seconds": dt_s, "gcell_s": cells * n / dt_s / 1e9, "steps": n, **seen}
out["slowdown"] = out["with_dumps"]["seconds"] / out["no_dumps"]["seconds"]
out["dump_gb_per_s"] = out["with_dumps"]["bytes"] / 1e9 / out["with_dumps"]["seconds"]
out["compute_s_between_dumps"] = out["no_dumps"]["seconds"] / steps * args.rate
out["bytes_per_dump"] = 6 * cells * 8
print(json.dumps(out))
