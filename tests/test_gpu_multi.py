"""z-slab decomposition with one process per slab: every slab's planes must equal the single-domain
oracle bit for bit, for the operator-level calls, for fdtd_run (halo traffic overlapped with the
interior planes) and for the dump variables.  Two ways of wiring the slabs are covered: peer memory
(CUDA IPC + sequence flags, fdtd_ctx_peer_connect) -- which also works when the slabs share one GPU,
so these tests run on a 1-GPU box -- and NCCL (fdtd_ctx_comm_init), which needs one GPU per slab."""
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


def _wire(F, dist, ctx, rank, world, transport):
    if transport == "nccl":
        box = [F.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0])
    else:
        blobs = [None] * world
        dist.all_gather_object(blobs, ctx.peer_export())
        ctx.peer_connect(blobs)


def _worker(rank, world, port, transport, cases, ret):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import fdtd_b200 as F
    import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        device = rank % torch.cuda.device_count()
        torch.cuda.set_device(device)
        o = O.restatement()
        bad = []
        for case, (dims, mode, steps, variant) in enumerate(cases):
            if world > dims[2]:
                continue
            nx, ny, nz = dims
            args = ((nx + .5) * DX, (ny + .5) * DX, (nz + .5) * DX, DX, DT, 1e-9, 1, mode)
            q, p = O.make_params(*args), F.make_params(*args)
            assert q.dims() == dims == p.dims()
            f = O.alloc_fields(*dims, rng=np.random.default_rng(5))
            init = {k[0].upper() + k[1:]: v.copy() for k, v in f.items()}

            def compare(ctx, what):
                got = ctx.download({k: np.zeros_like(v) for k, v in init.items()})
                k0, k1 = ctx.k0, ctx.k1
                top = 1 if rank == world - 1 else 0
                for name in ("Ez", "Hx", "Hy"):
                    if not np.array_equal(got[name][k0:k1].view(np.uint64), f[name.lower()][k0:k1].view(np.uint64)):
                        bad.append((case, what, name))
                for name in ("Ex", "Ey", "Hz"):
                    if not np.array_equal(got[name][k0:k1 + top].view(np.uint64),
                                          f[name.lower()][k0:k1 + top].view(np.uint64)):
                        bad.append((case, what, name))

            with F.Context(p, device=device, rank=rank, nranks=world) as ctx:
                for k, v in variant.items():
                    ctx.set_option(k, v)
                _wire(F, dist, ctx, rank, world, transport)
                assert ctx.get_option("transport") == (1 if transport == "nccl" else 3)
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
                # fused / overlapped path, queued as several runs without a sync in between
                t2 = ctx.run(steps - 2, t)
                t2 = ctx.run(1, t2)
                t2 = ctx.run(1, t2)
                t3 = o.run(q, f, steps, t)
                assert t2 == t3
                compare(ctx, "run")
                for v in range(6):
                    want = o.aggregate(q, f, v)[ctx.k0:ctx.k1]
                    if not np.array_equal(ctx.aggregate(v).view(np.uint64), want.view(np.uint64)):
                        bad.append((case, "aggregate", v))
                sums = ctx.checksum()
                # fdtd_run_hosted on slabs: host arrays of my planes in, 5 steps, host arrays out.  With peer memory
                # and the two-step kernel the slabs' wavefronts mesh (neighbours sweep in opposite directions);
                # everything else takes upload + run + download.
                ref = {k.lower(): v.copy() for k, v in init.items()}
                o.run(q, ref, 5)
                top = 1 if rank == world - 1 else 0
                k0, k1 = ctx.k0, ctx.k1
                for chunk in (2, 3):
                    ctx.set_option("host_chunk", chunk)
                    mine = {}
                    for name, arr in init.items():
                        hi = k1 + top if name in ("Ex", "Ey", "Hz") else k1
                        mine[name] = arr[k0:hi].copy()      # (a leading-axis slice is a view: the call works in place)
                    t_h = ctx.run_hosted(mine, 5, 0.0)
                    for name, arr in mine.items():
                        hi = k1 + top if name in ("Ex", "Ey", "Hz") else k1
                        if not np.array_equal(arr.view(np.uint64), ref[name.lower()][k0:hi].view(np.uint64)):
                            bad.append((case, f"hosted chunk {chunk}", name))
                    # the device state and its halos are current too: one more step from there
                    t_h = ctx.run(1, t_h)
                    ref1 = {k: v.copy() for k, v in ref.items()}
                    o.run(q, ref1, 1, 5 * DT)
                    got = ctx.download({k: np.zeros_like(v) for k, v in init.items()})
                    for name in ("Ez", "Hx", "Hy"):
                        if not np.array_equal(got[name][k0:k1].view(np.uint64), ref1[name.lower()][k0:k1].view(np.uint64)):
                            bad.append((case, f"hosted chunk {chunk} + run", name))
                dist.barrier()   # nobody unmaps a neighbour's memory while it may still be written
            allsums = [None] * world
            dist.all_gather_object(allsums, sums)
            total = [sum(s[a] for s in allsums) % (1 << 64) for a in range(6)]
            if total != F.checksum_host({k[0].upper() + k[1:]: v for k, v in f.items()}):
                bad.append((case, "checksum", total))
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
         ((70, 21, 8), 0, 6, dict(kernel=3, strip=1, kchunk=100, warps_x=1, warps_y=4, stages=4)),
         ((64, 21, 11), 1, 6, dict()),      # defaults; the last block in x holds the single column i = I
         # two steps per sweep: two halo planes each way, odd and even step counts, slabs of 2..3 planes
         ((37, 28, 23), 1, 7, dict(kernel=4, kchunk=8)),
         ((70, 21, 8), 0, 6, dict(kernel=4, kchunk=3, warps_y=12, stages=2)),
         ((33, 40, 9), 1, 8, dict(kernel=4, kchunk=1000, warps_y=16)),
         ((40, 33, 4), 1, 5, dict(kernel=4)),   # slabs too thin for the wide halos: single-step sweeps
         # in place on the rolling window (rings rotate in lockstep on every rank; uneven slabs have different rings)
         ((37, 28, 23), 1, 7, dict(kernel=4, rolling=1, kchunk=5)),
         ((33, 40, 9), 0, 6, dict(kernel=4, rolling=1))]


if os.environ.get("FDTD_MULTI_QUICK") == "1":  # a short list for expensive many-GPU boxes
    CASES = [CASES[0], CASES[4], CASES[8], CASES[9], CASES[11], CASES[12]]


def _spawn(target, world, *args):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port) + args + (ret,)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=600)
        assert pr.exitcode == 0
    return dict(ret)


def _layouts():
    """(world, transport) pairs this box can run: peer memory also when the slabs share GPUs (2..4 slabs on
    any box, 8 where 8 GPUs exist); NCCL only with one GPU per slab"""
    n = torch.cuda.device_count()
    out = [(w, "peer") for w in (2, 3, 4)] + ([(8, "peer")] if n >= 8 else [])
    return out + [(w, "nccl") for w in (2, 3, 4, 8) if n >= w]


@pytest.mark.parametrize("world,transport", _layouts() or [(2, "peer")])
def test_slabs_match_single_domain(world, transport):
    """all CASES in one set of `world` processes (spawning costs more than the cases)"""
    assert _spawn(_worker, world, transport, CASES) == {r: [] for r in range(world)}, f"world={world}"


def _propagate_worker(rank, world, port, transport, nums, ret):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import fdtd_b200 as F
    import oracle as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        device = rank % torch.cuda.device_count()
        torch.cuda.set_device(device)
        path = f"/tmp/fdtd_multi_prop_{os.getpid()}.txt"
        O.write_params(path, nums)
        p, q = F.load_parameters(path), O.restatement().load_parameters(path)
        o = O.restatement()
        f = O.alloc_fields(*q.dims())
        if q.mode == 0:
            o.set_initial_conditions(q, f)
        got = []
        with F.Context(p, device=device, rank=rank, nranks=world) as ctx:
            _wire(F, dist, ctx, rank, world, transport)
            if q.mode == 0:
                ctx.set_initial_conditions()
            steps, _ = ctx.propagate(on_begin=lambda it, dims, k0: got.append({"it": it, "dims": dims, "k0": k0, "vars": {}}),
                                     on_variable=lambda name, arr: got[-1]["vars"].__setitem__(name, arr))
            k0, k1 = ctx.k0, ctx.k1
            dist.barrier()
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


@pytest.mark.parametrize("transport", ["peer"] + (["nccl"] if torch.cuda.device_count() >= 2 else []))
@pytest.mark.parametrize("mode", [0, 1])
def test_slab_propagate_dumps_match_oracle(mode, transport):
    """fdtd_propagate on slabs: every rank dumps its own planes, same cadence, same contents
    (incl. the zone plane that averages with the upper neighbour's node plane, and aEy)."""
    world = 2 if transport == "nccl" else 3
    nums = ("0.021", "0.017", "0.013", "0.001", "0.0000000000006", "0.000000000012", "4", str(mode))
    assert _spawn(_propagate_worker, world, transport, nums) == {r: [] for r in range(world)}
