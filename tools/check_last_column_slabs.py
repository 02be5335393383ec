#!/usr/bin/env python
"""Slab check of the lone-last-column path of the default kernel: I a multiple of the tile width, every
slab count the box allows, fused default kernel, both modes; bit-exact against the oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_b200 as F  # noqa: E402
import oracle as O  # noqa: E402

o = O.restatement()
bad = 0
for world in sorted({1, 2, min(torch.cuda.device_count(), 4)}):
    for dims, mode in (((64, 21, 11), 1), ((32, 40, 9), 0), ((96, 17, 8), 1)):
        args = tuple((d + .5) * 1e-3 for d in dims) + (0.001, 6e-13, 1.2e-10, 2, mode)
        p, q = F.make_params(*args), O.make_params(*args)
        assert p.dims() == dims
        f = O.alloc_fields(*dims, rng=np.random.default_rng(3))
        with F.Group(p, world) as g:
            g.upload({k[0].upper() + k[1:]: v for k, v in f.items()})
            g.run(5, 0.0)
            o.run(q, f, 5)
            got = g.download()
        ok = all(np.array_equal(got[k[0].upper() + k[1:]].view(np.uint64), v.view(np.uint64)) for k, v in f.items())
        bad += not ok
        print(world, dims, mode, "ok" if ok else "MISMATCH", flush=True)
sys.exit(1 if bad else 0)
