"""The fused single-sweep step as an algorithm, on the CPU: tests/fused_numpy.py restates the kernels'
schedule (H and E of a plane in one pass, double-buffered state, chunk prologue, copies for PEC,
fused source) in numpy; it must reproduce the oracle bit for bit, alone and cut into z-slabs that
exchange halos over gloo with the plan the NCCL path uses (fdtd_b200.HALO_PLAN_FUSED)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from fused_numpy import fused_step

DX, DT = 0.001, 6e-13
MU, EPS = 1.25663706143591729538505735331180115367886775975E-6, 8.854E-12


def _setup(O, F, dims, mode):
    args = tuple((d + .5) * DX for d in dims) + (DX, DT, 1e-9, 1, mode)
    q, p = O.make_params(*args), F.make_params(*args)
    assert q.dims() == dims
    ch, ce = DT / (MU * DX), DT / (EPS * DX)
    plan = F.source_plan(p) if mode == 1 else None
    return q, p, ch, ce, plan


def _src(F, p, plan, t):
    if plan is None:
        return None
    ez, hx = F.source_values(p, plan, t)
    return (plan.i0, plan.i1, plan.j0, plan.j1, ez, hx)


@pytest.mark.parametrize("dims,mode,kchunk", [((23, 19, 12), 1, 4), ((23, 19, 12), 0, 5), ((9, 31, 7), 1, 1),
                                              ((30, 30, 1), 1, 8), ((1, 1, 5), 0, 2), ((17, 16, 9), 1, 100)])
def test_fused_schedule_equals_reference_loop(F, oracle, dims, mode, kchunk):
    o = oracle.restatement()
    q, p, ch, ce, plan = _setup(oracle, F, dims, mode)
    f = oracle.alloc_fields(*dims, rng=np.random.default_rng(11))
    a = {k: v.copy() for k, v in f.items()}
    b = {k: np.full_like(v, np.nan) for k, v in f.items()}   # every element must be written
    t = 0.0
    for _ in range(5):
        fused_step(a, b, dims, ch, ce, _src(F, p, plan, t), kchunk)
        a, b = b, a
        t += DT
    o.run(q, f, 5)
    for k in f:
        assert np.array_equal(a[k].view(np.uint64), f[k].view(np.uint64)), k


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _exchange(arrs, plan, rank, world, k0, k1):
    import torch
    to = plan["to"]
    send_k = k1 - 1 if plan["send_plane"] == "k1-1" else k0
    recv_k = k0 - 1 if plan["recv_plane"] == "k0-1" else k1
    reqs, bufs = [], []
    for name in plan["fields"]:
        arr = arrs[name.lower()]
        if 0 <= rank + to < world:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(arr[send_k])), rank + to))
        if 0 <= rank - to < world:
            buf = torch.empty(arr[recv_k].shape, dtype=torch.float64)
            reqs.append(dist.irecv(buf, rank - to))
            bufs.append((arr, recv_k, buf))
    for r in reqs:
        r.wait()
    for arr, idx, buf in bufs:
        arr[idx] = buf.numpy()


def _worker(rank, world, port, dims, mode, steps, ret):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fdtd_b200 as F
    import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = O.restatement()
        q, p, ch, ce, plan = _setup(O, F, dims, mode)
        want = O.alloc_fields(*dims, rng=np.random.default_rng(5))
        init = {k: v.copy() for k, v in want.items()}
        o.run(q, want, steps)
        k0, k1 = F.slab_range(dims[2], rank, world)
        # whole-cavity arrays, but everything outside this slab (and its halo planes) is poisoned
        a = {k: v.copy() for k, v in init.items()}
        for name, arr in a.items():
            node = name in ("ex", "ey", "hz")
            lo, hi = max(k0 - 1, 0), min(k1 + 1, arr.shape[0])
            arr[:lo] = np.nan
            arr[hi:] = np.nan
        b = {k: np.full_like(v, np.nan) for k, v in a.items()}
        t = 0.0
        for _ in range(steps):
            src = None
            # the source lives on the slab that holds k = 0 -- and the slab that starts at k = 1 needs the
            # amplitudes too: it recomputes H of plane k = 0 (the plane below it), source included
            if plan is not None and k0 <= 1:
                ez, hx = F.source_values(p, plan, t)
                src = (plan.i0, plan.i1, plan.j0, plan.j1, ez, hx)
            fused_step(a, b, dims, ch, ce, src, kchunk=3, klo=k0, khi=k1)
            a, b = b, a
            _exchange(a, F.HALO_PLAN_FUSED["after_step_up"], rank, world, k0, k1)
            _exchange(a, F.HALO_PLAN_FUSED["after_step_down"], rank, world, k0, k1)
            t += DT
        top = 1 if rank == world - 1 else 0
        bad = []
        for name in ("ez", "hx", "hy"):
            if not np.array_equal(a[name][k0:k1].view(np.uint64), want[name][k0:k1].view(np.uint64)):
                bad.append(name)
        for name in ("ex", "ey", "hz"):
            if not np.array_equal(a[name][k0:k1 + top].view(np.uint64), want[name][k0:k1 + top].view(np.uint64)):
                bad.append(name)
        ret[rank] = bad
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims,mode,steps", [(2, (23, 19, 12), 1, 6), (3, (21, 17, 10), 1, 5),
                                                    (2, (12, 14, 9), 0, 5), (4, (12, 14, 4), 1, 5)])
def test_fused_slabs_with_fused_halo_plan(world, dims, mode, steps):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, mode, steps, ret)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=180)
        assert pr.exitcode == 0
    assert dict(ret) == {r: [] for r in range(world)}
