"""All slabs from ONE process and ONE thread (fdtd_group_*, ncclCommInitAll): same results as the
single-domain oracle, bit for bit -- stepping, dumps, and the C host program with FDTD_B200_GPUS."""
import hashlib
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, bits_equal, to_oracle_params, upper
from pdb_reader import SiloFile

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "fdtd-maxwell-microwave-oven_b200", "microwave")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ngpu():
    return torch.cuda.device_count()


def layouts(worlds=(2, 3, 4, 8)):
    """(world, devices, transport): slabs share GPUs when the box has fewer GPUs than slabs (peer copies
    of a group work on one device); NCCL needs one GPU per slab."""
    out = []
    for world in worlds:
        out.append((world, [r % ngpu() for r in range(world)], "peer"))
        if world <= ngpu():
            out.append((world, list(range(world)), "nccl"))
    return out


VARIANTS = [dict(kernel=4, kchunk=4, warps_y=8, stages=2),
            dict(kernel=4, rolling=1, kchunk=3),
            dict(kernel=3, strip=2, kchunk=4, warps_x=2, warps_y=2, stages=3),
            dict(kernel=2, strip=1, kchunk=32, warps_x=1, warps_y=4),
            dict(kernel=1, strip=2, kchunk=3, warps_x=2, warps_y=2),
            dict(kernel=0)]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("mode", [0, 1])
def test_group_run_matches_oracle(F, oracle, variant, mode):
    o = oracle.restatement()
    for world, devices, transport in layouts():
        dims = (37, 28, 23)
        args = tuple((d + .5) * 1e-3 for d in dims) + (0.001, 6e-13, 1.2e-10, 2, mode)
        p, q = F.make_params(*args), oracle.make_params(*args)
        assert p.dims() == dims
        f = oracle.alloc_fields(*dims, rng=np.random.default_rng(17))
        with F.Group(p, world, devices=devices, transport=transport) as g:
            for k, v in variant.items():
                g.set_option(k, v)
            g.upload(upper(f))
            t = g.run(4, 0.0)
            t = g.run(3, t)
            t_cpu = o.run(q, f, 7)
            assert t == t_cpu
            got = g.download()
            for k, want in f.items():
                assert bits_equal(got[k[0].upper() + k[1:]], want), (world, transport, k)
            sums = [s.checksum() for s in g.slabs]
            total = [sum(s[a] for s in sums) % (1 << 64) for a in range(6)]
            assert total == F.checksum_host(upper(f))
            for v in range(6):
                assert bits_equal(g.aggregate(v), o.aggregate(q, f, v)), (world, v)
            _, h_ref = o.energy(q, f)
            e_gpu, h_gpu = g.energy(as_coded=False)
            ex, ey, ez = f["ex"], f["ey"], f["ez"]
            mex = (ex[:-1, :-1, :] + ex[1:, :-1, :] + ex[:-1, 1:, :] + ex[1:, 1:, :]) / 4.
            mey = (ey[:-1, :, :-1] + ey[:-1, :, 1:] + ey[1:, :, :-1] + ey[1:, :, 1:]) / 4.
            mez = (ez[:, :-1, :-1] + ez[:, 1:, :-1] + ez[:, :-1, 1:] + ez[:, 1:, 1:]) / 4.
            e_ref = ((mex ** 2).sum() + (mey ** 2).sum() + (mez ** 2).sum()) * q.spatial_step ** 3 * 8.854e-12 / 2.
            assert abs(e_gpu - e_ref) <= 1e-11 * abs(e_ref) and abs(h_gpu - h_ref) <= 1e-12 * abs(h_ref)
            with pytest.raises(F.FdtdError) as err:      # a slab of a group cannot exchange halos on its own
                g.slabs[0].aggregate(0)
            assert err.value.code == -6


@pytest.mark.parametrize("mode", [0, 1])
def test_group_propagate_matches_reference_dumps(F, golden, tmp_path, mode):
    g_ = golden["propagate_tiny"][f"mode{mode}"]
    path = tmp_path / "p.txt"
    path.write_text("\n".join(g_["params"]))
    p = F.load_parameters(path)
    nx, ny, nz = p.dims()
    logs = {}
    with F.Group(p, 2, devices=[0, 1 % ngpu()]) as g:
        if mode == 0:
            g.set_initial_conditions()
        steps, _ = g.propagate(
            on_begin=lambda r, it, dims, k0: logs.setdefault(it, {}).setdefault(r, {"k0": k0, "dims": dims, "vars": {}}),
            on_variable=lambda r, name, arr: logs[max(i for i in logs if r in logs[i])][r]["vars"].__setitem__(name, arr))
        assert steps == g_["steps"]
        out = g.download()
    its = sorted(logs)
    assert ["r/result%04d.silo" % i for i in its] == [d["file"] for d in g_["dumps"]]
    for it, want in zip(its, g_["dumps"]):
        slabs = [logs[it][r] for r in sorted(logs[it])]
        assert [s["k0"] for s in slabs] == [g.slabs[r].k0 for r in range(2)]
        for name, dig in want["vars"].items():
            whole = np.concatenate([s["vars"][name] for s in slabs])
            assert whole.size == nx * ny * nz and digest(whole) == dig, (it, name)
    for k, want in g_["final_sha256"].items():
        assert digest(out[k[0].upper() + k[1:]]) == want, k


def test_microwave_with_two_gpus(F, golden, tmp_path):
    """FDTD_B200_GPUS=2 ./microwave params.txt : one process, two slabs, bricks that add up."""
    g_ = golden["propagate_tiny"]["mode1"]
    (tmp_path / "params.txt").write_text("\n".join(g_["params"]))
    (tmp_path / "r").mkdir()
    r = subprocess.run([EXE, "params.txt"], cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, FDTD_B200_GPUS="2", FDTD_B200_DEVICES=f"0,{1 % ngpu()}",
                                FDTD_B200_SINK="raw"))
    assert r.returncode == 0, r.stderr
    nx, ny, nz = g_["grid"]
    k = [F.slab_range(nz, rr, 2) for rr in range(2)]
    names = ["ex", "ey", "ez", "hx", "hy", "hz"]
    for d in g_["dumps"]:
        base = os.path.basename(d["file"]).replace(".silo", "")
        parts = [np.fromfile(tmp_path / "r" / f"{base}.slab{rr}.raw") for rr in range(2)]
        for v, name in enumerate(names):
            whole = np.concatenate([parts[rr][v * nx * ny * (k[rr][1] - k[rr][0]):(v + 1) * nx * ny * (k[rr][1] - k[rr][0])]
                                    for rr in range(2)])
            assert digest(whole) == d["vars"][name], (base, name)


@pytest.mark.parametrize("dims,mode", [((64, 21, 11), 1), ((32, 40, 9), 0), ((96, 17, 8), 1)])
def test_group_default_kernel_when_the_last_block_holds_one_column(F, oracle, dims, mode):
    """I a multiple of the tile width: the default kernel's last block in x holds the single column
    i = I and takes the direct path (last_column_sweep); slabs must still match the oracle.
    (tools/check_last_column_slabs.py is the same check as a script.)"""
    o = oracle.restatement()
    for world, devices, transport in layouts((2, 4)):
        args = tuple((d + .5) * 1e-3 for d in dims) + (0.001, 6e-13, 1.2e-10, 2, mode)
        p, q = F.make_params(*args), oracle.make_params(*args)
        assert p.dims() == dims
        f = oracle.alloc_fields(*dims, rng=np.random.default_rng(3))
        with F.Group(p, world, devices=devices, transport=transport) as g:
            g.upload(upper(f))
            g.run(5, 0.0)
            o.run(q, f, 5)
            got = g.download()
        for k, want in f.items():
            assert bits_equal(got[k[0].upper() + k[1:]], want), (world, k)


@pytest.mark.parametrize("nz", [6, 13])
@pytest.mark.parametrize("kernel", [4, 3, 1, "rolling"])
def test_many_short_runs_back_to_back_on_thin_slabs(F, oracle, kernel, nz):
    """Queue many short runs without synchronising in between, on slabs of one or two planes: every
    run must see the halos the previous one sent (the exchange of run n is still in flight when the
    host queues run n + 1)."""
    o = oracle.restatement()
    dims = (70, 45, nz)
    args = tuple((d + .5) * 1e-3 for d in dims) + (0.001, 6e-13, 1.2e-10, 2, 1)
    p, q = F.make_params(*args), oracle.make_params(*args)
    assert p.dims() == dims
    for world, devices, transport in layouts((4,)):
        f = oracle.alloc_fields(*dims, rng=np.random.default_rng(11))
        with F.Group(p, world, devices=devices, transport=transport) as g:
            if kernel == "rolling":
                g.set_option("kernel", 4)
                g.set_option("rolling", 1)
            else:
                g.set_option("kernel", kernel)
            g.upload(upper(f))
            t, total = 0.0, 0
            for n in range(20):
                t = g.run(1 + n % 3, t)          # 1, 2, 3, 1, ... steps: single sweeps and two-step sweeps mixed
                total += 1 + n % 3
            t_cpu = o.run(q, f, total)
            assert t == t_cpu
            got = g.download()
        for k, want in f.items():
            assert bits_equal(got[k[0].upper() + k[1:]], want), (world, transport, k)


def test_microwave_two_slabs_write_silo_multiblock(F, golden, tmp_path):
    """FDTD_B200_GPUS=2 with the default sink: one block file per slab and a root file with multimesh /
    multivar objects naming the blocks; the blocks' data add up to the reference's dump."""
    g_ = golden["propagate_tiny"]["mode0"]
    (tmp_path / "params.txt").write_text("\n".join(g_["params"]))
    (tmp_path / "r").mkdir()
    r = subprocess.run([EXE, "params.txt"], cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, FDTD_B200_GPUS="2", FDTD_B200_DEVICES=f"0,{1 % ngpu()}"))
    assert r.returncode == 0, r.stderr
    nx, ny, nz = g_["grid"]
    dx = float(g_["params"][3])
    names = ["ex", "ey", "ez", "hx", "hy", "hz", "aEy", "aHx", "aHz"]
    for d in g_["dumps"]:
        root = SiloFile(tmp_path / d["file"])
        base = os.path.basename(d["file"]).replace(".silo", "")
        mm = root.object("mesh")
        assert mm["_type"] == "multimesh" and mm["nblocks"] == 2
        assert mm["meshnames"] == f";{base}.slab0.silo:/mesh;{base}.slab1.silo:/mesh"
        assert root.objects() == ["mesh"] + names + ["vecs"]
        blocks = [SiloFile(tmp_path / "r" / f"{base}.slab{rr}.silo") for rr in range(2)]
        z = np.concatenate([b.object("mesh")["coord2"] for b in blocks])
        k = [F.slab_range(nz, rr, 2) for rr in range(2)]
        assert np.array_equal(z, np.concatenate([np.arange(a, b + 1) * dx for a, b in k]))
        for name in names:
            assert root.object(name)["_type"] == "multivar"
            whole = np.concatenate([b.object(name)["value0"] for b in blocks])
            assert digest(whole) == d["vars"][name], (base, name)
