"""Multi-rank host logic on CPU: world_size 2 and 3 over gloo.

The z-slab partition (fdtd_slab_range) and the halo plan (fdtd_b200.HALO_PLAN, the plan
exchange_h / exchange_e implement over NCCL) are run with the CPU oracle standing in for the
per-slab kernels; the owned planes of every rank must equal the single-domain result bit for bit.
Each rank steps a local cavity that covers its slab plus one plane either side; whatever the
local PEC treatment does to those outer planes is overwritten by the halo exchange before use.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

DX, DT = 0.001, 6e-13


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _exchange(local, plan, rank, world, a, k0, k1):
    """one HALO_PLAN entry over gloo; local arrays are indexed by global plane minus a"""
    to = plan["to"]
    send_k = k1 - 1 if plan["send_plane"] == "k1-1" else k0
    recv_k = k0 - 1 if plan["recv_plane"] == "k0-1" else k1
    reqs, bufs = [], []
    for name in plan["fields"]:
        arr = local[name.lower()]
        if 0 <= rank + to < world:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(arr[send_k - a])), rank + to))
        if 0 <= rank - to < world:
            buf = torch.empty(arr[recv_k - a].shape, dtype=torch.float64)
            reqs.append(dist.irecv(buf, rank - to))
            bufs.append((arr, recv_k - a, buf))
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
        nx, ny, nz = dims
        gp = O.make_params((nx + .5) * DX, (ny + .5) * DX, (nz + .5) * DX, DX, DT, 1e-9, 1, mode)
        assert gp.dims() == dims
        want = O.alloc_fields(*dims, rng=np.random.default_rng(5))
        init = {k: v.copy() for k, v in want.items()}
        o.run(gp, want, steps)                                   # single-domain truth

        k0, k1 = F.slab_range(nz, rank, world)
        a, b = max(k0 - 1, 0), min(k1 + 1, nz)
        lp = O.make_params((nx + .5) * DX, (ny + .5) * DX, (b - a + .5) * DX, DX, DT, 1e-9, 1, mode if a == 0 else 0)
        assert lp.dims() == (nx, ny, b - a)
        node, cell = ("ex", "ey", "hz"), ("ez", "hx", "hy")
        local = {k: np.ascontiguousarray(init[k][a:b + 1]) for k in node}
        local.update({k: np.ascontiguousarray(init[k][a:b]) for k in cell})
        t = 0.0
        for _ in range(steps):
            if lp.mode == 1:
                o.set_source(lp, local, t)
            o.update_h(lp, local)
            _exchange(local, F.HALO_PLAN["after_H"], rank, world, a, k0, k1)
            if lp.mode == 1:
                o.set_source(lp, local, t)
            o.update_e(lp, local)
            _exchange(local, F.HALO_PLAN["after_E"], rank, world, a, k0, k1)
            t += DT
        top = 1 if rank == world - 1 else 0
        bad = []
        for k in cell:
            if not np.array_equal(local[k][k0 - a:k1 - a].view(np.uint64), want[k][k0:k1].view(np.uint64)):
                bad.append(k)
        for k in node:
            if not np.array_equal(local[k][k0 - a:k1 - a + top].view(np.uint64), want[k][k0:k1 + top].view(np.uint64)):
                bad.append(k)
        ret[rank] = bad
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dims,mode,steps", [(2, (23, 19, 12), 1, 6), (2, (23, 19, 11), 0, 5),
                                                    (3, (21, 17, 10), 1, 5)])
def test_slab_plan_reproduces_single_domain(world, dims, mode, steps):
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
