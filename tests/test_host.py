"""Host half of the product (no GPU): the C ABI loads and exports every declared symbol, and the
host-side functions of the path agree with the oracle bit for bit."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, bits_equal, to_oracle_params


def test_library_exports_every_declared_symbol(F):
    header = open(os.path.join(ROOT, "include", "fdtd_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fdtd_[a-z_A-Z0-9]+)\s*\(", header))
    declared -= {"fdtd_dump_sink"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(F.lib, name), f"{name} is declared in include/fdtd_b200.h but not exported"
    assert set(F.SIGNATURES) == declared
    assert F.lib.fdtd_abi_version() == 1


def test_no_silent_cpu_path(F):
    """Without a CUDA device every device entry point must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = F.make_params(0.05, 0.05, 0.05, 0.001, 6e-13, 1.2e-10, 2, 1)
    with pytest.raises(F.FdtdError) as e:
        F.Context(p)
    assert e.value.code == -3  # FDTD_E_CUDA


def test_load_parameters_matches_oracle(F, oracle, tmp_path):
    o = oracle.restatement()
    cases = [oracle.STOCK_PARAMS,
             ("0.05", "0.04", "0.03", "0.001", "0.0000000000006", "0.00000000012", "2", "1"),
             ("1.024", "1.024", "1.024", "0.001", "6e-13", "6e-11", "50", "1"),
             ("0.256", "0.256", "0.256", "0.001", "0.0000000000006", "0.0000000006", "1000000", "1")]
    for nums in cases:
        path = oracle.write_params(tmp_path / "p.txt", nums)
        a, b = F.load_parameters(path), o.load_parameters(path)
        assert a.dims() == b.dims()
        for name in ("length", "width", "height", "spatial_step", "time_step", "simulation_time",
                     "sampling_rate", "mode"):
            assert getattr(a, name) == getattr(b, name), name
        assert F.step_count(a) == oracle.step_count(b)
        want = oracle.field_shapes(*b.dims())
        assert F.field_sizes(a) == [int(np.prod(want[k])) for k in oracle.FIELD_NAMES]
    with pytest.raises(F.FdtdError) as e:
        F.load_parameters(tmp_path / "missing.txt")
    assert e.value.code == -2 and "Unable to open parameters file!" in str(e.value)


@pytest.mark.parametrize("dims", [(0.05, 0.05, 0.05), (0.05, 0.04, 0.03), (0.256, 0.256, 0.256),
                                  (0.037, 0.029, 0.023)])
def test_source_values_match_oracle_set_source(F, oracle, dims):
    """fdtd_source_values == what set_source (main.c:748,751) writes, at several times."""
    o = oracle.restatement()
    p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, 1)
    q = to_oracle_params(oracle, p)
    plan = F.source_plan(p)
    assert (plan.i0, plan.i1, plan.j0, plan.j1) == oracle.source_bounds(q)
    assert plan.z_te == oracle.source_zte(q)
    if max(q.dims()) > 64:
        q = oracle.make_params(*dims[:2], 0.002, 0.001, 6e-13, 1.2e-10, 2, 1)  # thin z: same patch
    for t in (0.0, 6e-13, 7 * 6e-13, 1.19994e-10):
        f = oracle.alloc_fields(*q.dims())
        o.set_source(q, f, t)
        ez, hx = F.source_values(p, plan, t)
        for s in range(plan.i1 - plan.i0):
            col_e = f["ez"][0, plan.j0:plan.j1, plan.i0 + s]
            col_h = f["hx"][0, plan.j0:plan.j1, plan.i0 + s]
            assert bits_equal(col_e, np.full_like(col_e, ez[s]))
            assert bits_equal(col_h, np.full_like(col_h, hx[s]))


def test_initial_conditions_match_oracle(F, oracle):
    o = oracle.restatement()
    for dims in ((0.05, 0.05, 0.05), (0.033, 0.017, 0.009)):
        p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, 0)
        q = to_oracle_params(oracle, p)
        f = oracle.alloc_fields(*q.dims())
        o.set_initial_conditions(q, f)
        assert bits_equal(F.initial_conditions_host(p), f["ey"])


def test_slab_ranges_partition_the_planes(F):
    for maxk in (1, 7, 50, 256, 1024, 2048):
        for n in (1, 2, 3, 4, 8):
            if n > maxk:
                with pytest.raises(F.FdtdError):
                    r = [F.slab_range(maxk, k, n) for k in range(n)]
                    if any(a == b for a, b in r):
                        raise F.FdtdError(-1, "empty slab")
                continue
            r = [F.slab_range(maxk, k, n) for k in range(n)]
            assert r[0][0] == 0 and r[-1][1] == maxk
            assert all(r[k][1] == r[k + 1][0] for k in range(n - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(F.FdtdError):
        F.slab_range(10, 3, 3)


def test_pattern_and_checksum_mirrors_are_consistent(F):
    p = F.make_params(0.009, 0.007, 0.005, 0.001, 6e-13, 1.2e-10, 2, 0)
    a = F.pattern_host(p, 42)
    b = F.pattern_host(p, 43)
    assert all(abs(v).max() < 1.0 for v in a.values())
    assert F.checksum_host(a) != F.checksum_host(b)
    c = {k: v.copy() for k, v in a.items()}
    c["Hy"][1, 2, 3], c["Hy"][1, 2, 2] = c["Hy"][1, 2, 2], c["Hy"][1, 2, 3]  # a swap must show
    assert F.checksum_host(c)[4] != F.checksum_host(a)[4]
    assert F.checksum_host(c)[:4] == F.checksum_host(a)[:4]


def test_host_program_cli_errors_without_gpu(F, tmp_path):
    """The C host program's argument / file / parameter checks (main.c:811-821) come before any device
    work, so they behave the same on a box without a GPU: perror text, exit status 1."""
    import subprocess
    exe = os.path.join(ROOT, "fdtd-maxwell-microwave-oven_b200", "microwave")
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1
    assert r.stdout.splitlines() == ["Welcome into our microwave oven eletrico-magnetic field simulator! "]
    assert r.stderr.startswith("This program needs 1 argument: the parameters file (.txt). Eg.: ./microwave param.txt")
    r = subprocess.run([exe, "a", "b"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "This program needs 1 argument" in r.stderr
    r = subprocess.run([exe, "missing.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1
    assert r.stdout.splitlines()[-1] == "Loading the parameters..."
    assert r.stderr.strip() == "Unable to open parameters file!: No such file or directory"
    (tmp_path / "p.txt").write_text("0.012\n0.011\n0.010\n0.001\n0.0000000000006\n0.0000000000001\n3\n1")
    r = subprocess.run([exe, "p.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("The time step must be lower than the simulation time!")


def test_host_helpers_match_oracle_on_random_parameters(F, oracle, tmp_path):
    """300 random parameter files: the grid derivation through float, the step count and the source
    plan of the product agree with the oracle exactly (these decide everything downstream)."""
    rng = np.random.default_rng(7)
    o = oracle.restatement()
    for _ in range(300):
        dims = [f"{rng.uniform(0.003, 0.3):.{int(rng.integers(2, 7))}f}" for _ in range(3)]
        dx = f"{rng.choice([0.001, 0.0005, 0.002, 0.00125]):g}"
        dt = f"{rng.uniform(2e-13, 9e-13):.3e}"
        sim = f"{rng.uniform(5e-12, 3e-10):.4e}"
        nums = (*dims, dx, dt, sim, str(int(rng.integers(1, 60))), str(int(rng.integers(0, 2))))
        path = oracle.write_params(tmp_path / "p.txt", nums)
        a, b = F.load_parameters(path), o.load_parameters(path)
        assert a.dims() == b.dims(), nums
        assert F.step_count(a) == oracle.step_count(b), nums
        plan = F.source_plan(a)
        assert (plan.i0, plan.i1, plan.j0, plan.j1) == oracle.source_bounds(b), nums
        assert plan.z_te == oracle.source_zte(b) or (np.isnan(plan.z_te) and np.isnan(oracle.source_zte(b))), nums


def test_constants_of_the_header_match_the_glue(F):
    """FDTD_PEER_BLOB_BYTES in include/fdtd_b200.h is what the ctypes glue allocates; the NUMA query works
    without a device (one node or none reported for the GPU is fine: the allocation then uses the default policy)."""
    import re
    from conftest import ROOT
    text = open(os.path.join(ROOT, "include", "fdtd_b200.h")).read()
    assert int(re.search(r"#define FDTD_PEER_BLOB_BYTES (\d+)", text).group(1)) == F.PEER_BLOB_BYTES
    info = F.host_numa_info(0)
    assert info["nodes"] >= 1 and info["node_of_device"] >= -1


def test_multi_slab_calls_fail_loudly_without_a_device(F):
    """no CPU path behind the slab / group entry points either"""
    p = F.make_params(0.02, 0.02, 0.02, 0.001, 6e-13, 1e-10, 2, 1)
    if __import__("torch").cuda.is_available():
        pytest.skip("needs a box without a GPU")
    with pytest.raises(F.FdtdError) as e:
        F.Context(p, device=0, rank=1, nranks=2)
    assert e.value.code == -3
    with pytest.raises(F.FdtdError) as e:
        F.Group(p, 2, devices=[0, 0])
    assert e.value.code == -3
