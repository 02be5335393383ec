#!/usr/bin/env python
"""Kernel-variant sweep on one GPU: Gcell-updates/s and per-kernel GB/s for each launch shape.
    python tools/sweep.py [--nz 512] [--steps 5]  > gpurun_out/sweep.txt
"""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdtd_b200 as F  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--nz", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--quick", action="store_true")
ap.add_argument("--small", action="store_true")
ap.add_argument("--big", action="store_true")
ap.add_argument("--fine", action="store_true")
ap.add_argument("--tma", action="store_true")
ap.add_argument("--prefetch", action="store_true")
ap.add_argument("--step2", action="store_true", help="the two-steps-per-sweep kernel (4)")
ap.add_argument("--only", type=int, default=None, help="only this kernel id (plus the one-thread-per-cell reference)")
args = ap.parse_args()

p = F.make_params(args.n * 1e-3, args.n * 1e-3, args.nz * 1e-3, 1e-3, 6e-13, 1e-9, 1 << 30, 1)
assert p.dims() == (args.n, args.n, args.nz), p.dims()
cells = args.n * args.n * args.nz
variants = [dict(kernel=0)]
shapes = [(1, 8), (2, 4), (4, 2), (8, 1), (1, 4), (2, 2), (4, 1)]
for strip, kchunk, (wx, wy) in itertools.product((1, 2, 4), (8, 32, 128), shapes):
    if args.quick and (kchunk != 32 or (wx, wy) not in ((2, 4), (4, 2))):
        continue
    variants.append(dict(kernel=1, strip=strip, kchunk=kchunk, warps_x=wx, warps_y=wy))

for strip, kchunk, (wx, wy) in itertools.product((1, 2, 3, 4), (8, 32, 128), [(1, 4), (2, 2), (4, 1), (1, 2), (2, 1)]):
    if args.quick and kchunk != 32:
        continue
    variants.append(dict(kernel=2, strip=strip, kchunk=kchunk, warps_x=wx, warps_y=wy))
if args.prefetch:
    variants = [dict(kernel=0)]
    for strip, kchunk, pf, (wx, wy) in itertools.product((1, 2), (32, 128), (0, 1, 2, 3, 4, 6, 8), [(2, 2), (1, 4), (4, 1)]):
        variants.append(dict(kernel=2, strip=strip, kchunk=kchunk, prefetch=pf, warps_x=wx, warps_y=wy))
if args.tma:
    variants = [dict(kernel=0), dict(kernel=2, strip=1, kchunk=32, prefetch=3, warps_x=1, warps_y=4)]
    for strip, kchunk, stages, (wx, wy) in itertools.product((1, 2), (32, 128), (2, 3, 4, 6), [(1, 4), (1, 8), (2, 4), (2, 2), (4, 2), (4, 1), (8, 1)]):
        variants.append(dict(kernel=3, strip=strip, kchunk=kchunk, stages=stages, warps_x=wx, warps_y=wy))
if args.fine:
    variants = [dict(kernel=0)]
    for strip, kchunk, stages, (wx, wy) in itertools.product((2,), (16, 32, 64, 128), (2, 3, 4), [(4, 2), (4, 1), (2, 2), (2, 4), (6, 1), (3, 2)]):
        variants.append(dict(kernel=3, strip=strip, kchunk=kchunk, stages=stages, warps_x=wx, warps_y=wy))
    for kchunk, stages, (wx, wy) in itertools.product((32, 64), (3, 4, 6), [(4, 2), (2, 4), (1, 8), (4, 1)]):
        variants.append(dict(kernel=3, strip=1, kchunk=kchunk, stages=stages, warps_x=wx, warps_y=wy))
if args.big:
    variants = [dict(kernel=3, strip=2, kchunk=32, stages=3, warps_x=4, warps_y=2)]
    for kchunk, stages, (wx, wy) in itertools.product((32, 64), (2, 3), [(4, 4), (2, 8)]):
        variants.append(dict(kernel=3, strip=2, kchunk=kchunk, stages=stages, warps_x=wx, warps_y=wy))
    variants.append(dict(kernel=3, strip=2, kchunk=32, stages=3, warps_x=4, warps_y=2))
if args.small:
    variants = []
    for kchunk, (wx, wy) in itertools.product((4, 8, 16, 32, 64), [(4, 2), (2, 2), (2, 4), (1, 8)]):
        variants.append(dict(kernel=3, strip=2 if (wx, wy) != (1, 8) else 1, kchunk=kchunk, stages=3 if (wx, wy) != (1, 8) else 4, warps_x=wx, warps_y=wy))
    variants.append(dict(kernel=1, strip=2, kchunk=8, warps_x=2, warps_y=2))
if args.step2:
    variants = [dict(kernel=0), dict(kernel=3, strip=1, kchunk=32, stages=4, warps_x=1, warps_y=8)]
    for wy, stages, kchunk in itertools.product((8, 16), (2, 3, 4), (24, 32, 64)):
        variants.append(dict(kernel=4, warps_y=wy, stages=stages, kchunk=kchunk, persistent=0))
    for wy, stages, window in itertools.product((8, 16), (3, 4), (2, 8)):
        variants.append(dict(kernel=4, warps_y=wy, stages=stages, kchunk=32, persistent=1, window=window))
    variants.append(dict(kernel=4, warps_y=8, stages=3, kchunk=32, persistent=0))
if args.only is not None:
    variants = [v for v in variants if v["kernel"] == args.only or v["kernel"] == 0]

ref_sum = None
with F.Context(p) as ctx:
    for v in variants:
        for k, val in v.items():
            ctx.set_option(k, val)
        try:
            ctx.run(1, 0.0)
        except F.FdtdError as e:
            print(json.dumps(dict(v, error=str(e))), flush=True)
            continue
        ctx.fill_test_pattern(1)
        t = ctx.run(2, 0.0)
        ctx.sync()
        t, total, h, e = ctx.run_timed(args.steps, t)
        t, total2, h, e = ctx.run_timed(args.steps, t)
        total = min(total, total2)
        s = ctx.checksum()
        if ref_sum is None:
            ref_sum = s
        ok = s == ref_sum
        gcs = cells * args.steps / (total * 1e-3) / 1e9
        hb = 72.0 * cells / (max(h, 1e-6) / args.steps * 1e-3) / 1e9
        eb = 72.0 * cells / (max(e, 1e-6) / args.steps * 1e-3) / 1e9
        print(json.dumps(dict(v, gcell_s=round(gcs, 2), h_gbs=round(hb), e_gbs=round(eb), same_bits=ok)), flush=True)
