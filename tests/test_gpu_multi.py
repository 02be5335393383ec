"""z-slab decomposition on real GPUs (one process per GPU, NCCL halos): every slab's planes must
equal the single-domain oracle bit for bit, for the operator-level calls, for fdtd_run (halo
traffic overlapped with the interior planes) and for the dump variables.  Skipped on a 1-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
DX, DT = 0.001, 6e-13


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dims, mode, steps, variant, ret):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import fdtd_b200 as F
    import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(rank)
        nx, ny, nz = dims
        args = ((nx + .5) * DX, (ny + .5) * DX, (nz + .5) * DX, DX, DT, 1e-9, 1, mode)
        q, p = O.make_params(*args), F.make_params(*args)
        assert q.dims() == dims == p.dims()
        o = O.restatement()
        f = O.alloc_fields(*dims, rng=np.random.default_rng(5))
        init = {k[0].upper() + k[1:]: v.copy() for k, v in f.items()}
        bad = []

        def compare(ctx, what):
            got = ctx.download({k: np.zeros_like(v) for k, v in init.items()})
            k0, k1 = ctx.k0, ctx.k1
            top = 1 if rank == world - 1 else 0
            for name in ("Ez", "Hx", "Hy"):
                if not np.array_equal(got[name][k0:k1].view(np.uint64), f[name.lower()][k0:k1].view(np.uint64)):
                    bad.append((what, name))
            for name in ("Ex", "Ey", "Hz"):
                if not np.array_equal(got[name][k0:k1 + top].view(np.uint64),
                                      f[name.lower()][k0:k1 + top].view(np.uint64)):
                    bad.append((what, name))

        with F.Context(p, device=rank, rank=rank, nranks=world) as ctx:
            for k, v in variant.items():
                ctx.set_option(k, v)
            box = [F.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            ctx.comm_init(box[0])
            ctx.upload(init)
            # operator level, like the reference's loop body
            t = 0.0
            for _ in range(2):
                if mode == 1:
                    ctx.set_source(t); o.set_source(q, f, t)
                ctx.update_H_field(); o.update_h(q, f)
                if mode == 1:
                    ctx.set_source(t); o.set_source(q, f, t)
                ctx.update_E_field(); o.update_e(q, f)
                t += DT
            compare(ctx, "operators")
            # fused / overlapped path
            t2 = ctx.run(steps, t)
            t3 = o.run(q, f, steps, t)
            assert t2 == t3
            compare(ctx, "run")
            for v in range(6):
                want = o.aggregate(q, f, v)[ctx.k0:ctx.k1]
                if not np.array_equal(ctx.aggregate(v).view(np.uint64), want.view(np.uint64)):
                    bad.append(("aggregate", v))
            sums = ctx.checksum()
        allsums = [None] * world
        dist.all_gather_object(allsums, sums)
        total = [sum(s[a] for s in allsums) % (1 << 64) for a in range(6)]
        if total != F.checksum_host({k[0].upper() + k[1:]: v for k, v in f.items()}):
            bad.append(("checksum", total))
        ret[rank] = bad
    finally:
        dist.destroy_process_group()


CASES = [((37, 28, 23), 1, 7, dict(kernel=1, strip=2, kchunk=8, warps_x=2, warps_y=2)),
         ((37, 28, 23), 1, 5, dict(kernel=0)),
         ((33, 40, 9), 0, 6, dict(kernel=1, strip=4, kchunk=3, warps_x=1, warps_y=4)),
         ((70, 21, 8), 1, 6, dict(kernel=1, strip=1, kchunk=100, warps_x=4, warps_y=2)),
         ((37, 28, 23), 1, 7, dict(kernel=2, strip=2, kchunk=8, warps_x=2, warps_y=2)),
         ((33, 40, 9), 0, 6, dict(kernel=2, strip=3, kchunk=2, warps_x=1, warps_y=4)),
         ((70, 21, 8), 1, 6, dict(kernel=2, strip=1, kchunk=100, warps_x=4, warps_y=1)),
         ((40, 33, 4), 1, 5, dict(kernel=2, strip=4, kchunk=16, warps_x=1, warps_y=2)),
         ((37, 28, 23), 1, 7, dict(kernel=3, strip=2, kchunk=8, warps_x=2, warps_y=2, stages=3)),
         ((70, 21, 8), 0, 6, dict(kernel=3, strip=1, kchunk=100, warps_x=1, warps_y=4, stages=4))]


if os.environ.get("FDTD_MULTI_QUICK") == "1":  # a short list for expensive many-GPU boxes
    CASES = [CASES[0], CASES[4], CASES[8], CASES[9]]


@pytest.mark.parametrize("dims,mode,steps,variant", CASES)
def test_slabs_match_single_domain(dims, mode, steps, variant):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    for world in sorted({2, min(ngpu, 4), min(ngpu, 8)}):
        if world > dims[2]:
            continue
        ctx = mp.get_context("spawn")
        ret = ctx.Manager().dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, dims, mode, steps, variant, ret))
                 for r in range(world)]
        for pr in procs:
            pr.start()
        for pr in procs:
            pr.join(timeout=300)
            assert pr.exitcode == 0
        assert dict(ret) == {r: [] for r in range(world)}, f"world={world}"


def _propagate_worker(rank, world, port, nums, ret):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import fdtd_b200 as F
    import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(rank)
        path = f"/tmp/fdtd_multi_prop_{os.getpid()}.txt"
        O.write_params(path, nums)
        p, q = F.load_parameters(path), O.restatement().load_parameters(path)
        o = O.restatement()
        f = O.alloc_fields(*q.dims())
        if q.mode == 0:
            o.set_initial_conditions(q, f)
        got = []
        with F.Context(p, device=rank, rank=rank, nranks=world) as ctx:
            box = [F.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            ctx.comm_init(box[0])
            if q.mode == 0:
                ctx.set_initial_conditions()
            steps, _ = ctx.propagate(on_begin=lambda it, dims, k0: got.append({"it": it, "dims": dims, "k0": k0, "vars": {}}),
                                     on_variable=lambda name, arr: got[-1]["vars"].__setitem__(name, arr))
            k0, k1 = ctx.k0, ctx.k1
        nx, ny, nz = q.dims()
        names = ["ex", "ey", "ez", "hx", "hy", "hz"]

        def expected(t_val):
            out = {n: o.aggregate(q, f, v)[k0:k1] for v, n in enumerate(names)}
            if q.mode == 0:
                v = o.validation_fields(q, f, t_val)
                out["aEy"] = o.aggregate(q, dict(f, ey=v["ey"]), 1)[k0:k1]
                out["aHx"], out["aHz"] = out["hx"], out["hz"]   # as coded, main.c:585-588
            return out

        bad = []
        want = [(1, expected(0.0))]
        t, it = 0.0, 1
        for _ in range(O.step_count(q)):
            o.run(q, f, 1, t)
            if it % q.sampling_rate == 0:
                want.append((it, expected(t)))
            t += q.time_step
            it += 1
        if steps != O.step_count(q) or [d["it"] for d in got] != [w[0] for w in want]:
            bad.append(("cadence", [d["it"] for d in got]))
        else:
            for d, (it_w, vars_w) in zip(got, want):
                if tuple(d["dims"]) != (nx, ny, k1 - k0) or d["k0"] != k0 or list(d["vars"]) != list(vars_w):
                    bad.append(("shape", d["it"]))
                    continue
                for name, arr in vars_w.items():
                    if not np.array_equal(d["vars"][name].view(np.uint64), arr.reshape(-1).view(np.uint64)):
                        bad.append((d["it"], name))
        ret[rank] = bad
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", [0, 1])
def test_slab_propagate_dumps_match_oracle(mode):
    """fdtd_propagate on slabs: every rank dumps its own planes, same cadence, same contents
    (incl. the zone plane that averages with the upper neighbour's node plane, and aEy)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    nums = ("0.021", "0.017", "0.013", "0.001", "0.0000000000006", "0.000000000012", "4", str(mode))
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_propagate_worker, args=(r, world, port, nums, ret)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=300)
        assert pr.exitcode == 0
    assert dict(ret) == {r: [] for r in range(world)}
