#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanity_small.py
Grids are chosen so that blocks touch every wall and the TMA boxes hang over every edge."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_b200 as F  # noqa: E402
import oracle as O  # noqa: E402

o = O.restatement()
bad = 0
for dims, mode in (((0.037, 0.029, 0.011), 1), ((0.07, 0.005, 0.006), 0), ((0.013, 0.075, 0.004), 1)):
    for variant in (dict(kernel=0), dict(kernel=1, strip=2, kchunk=3, warps_x=2, warps_y=2),
                    dict(kernel=1, strip=4, kchunk=32, warps_x=1, warps_y=4),
                    dict(kernel=2, strip=1, kchunk=4, warps_x=1, warps_y=4, prefetch=3),
                    dict(kernel=2, strip=3, kchunk=32, warps_x=2, warps_y=2, prefetch=2),
                    dict(kernel=3, strip=2, kchunk=4, warps_x=4, warps_y=2, stages=3),
                    dict(kernel=3, strip=1, kchunk=32, warps_x=1, warps_y=4, stages=2)):
        if mode == 1 and min(dims[:2]) < 0.012:
            continue
        p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, mode)
        q = O.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, mode)
        f = O.alloc_fields(*q.dims(), rng=np.random.default_rng(1))
        with F.Context(p) as ctx:
            for k, v in variant.items():
                ctx.set_option(k, v)
            ctx.upload({k[0].upper() + k[1:]: v for k, v in f.items()})
            ctx.run(3, 0.0)
            o.run(q, f, 3)
            got = ctx.download()
            for v in range(6):
                ctx.aggregate(v)
            ctx.energy()
            ctx.checksum()
            if mode == 0:
                ctx.validation_error(1e-12)
        ok = all(np.array_equal(got[k[0].upper() + k[1:]].view(np.uint64), v.view(np.uint64)) for k, v in f.items())
        bad += not ok
        print(q.dims(), mode, variant, "ok" if ok else "MISMATCH", flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
