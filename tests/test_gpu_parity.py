"""Parity of the CUDA path (through the C ABI, libfdtd_b200.so) against the CPU oracle and the
golden fixtures.  Bit-exact: the kernels use un-fused IEEE double in the reference's operand
order, so every comparison is on the raw 64-bit patterns (tolerance: zero)."""
import hashlib

import numpy as np
import pytest

from conftest import bits_equal, to_oracle_params, upper

pytestmark = pytest.mark.gpu


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assert_fields_equal(got, want, what=""):
    for k in want:
        g, w = got[k[0].upper() + k[1:]], want[k]
        if not bits_equal(g, w):
            bad = np.argwhere(g.view(np.uint64) != w.view(np.uint64))
            raise AssertionError(f"{what} {k}: {len(bad)} of {w.size} elements differ, first at "
                                 f"(k,j,i)={tuple(bad[0])}: got {g[tuple(bad[0])]!r} want {w[tuple(bad[0])]!r}")


def random_state(O, p, seed):
    return O.alloc_fields(*p.dims(), rng=np.random.default_rng(seed))


VARIANTS = [dict(kernel=0),
            dict(kernel=1, strip=1, kchunk=32, warps_x=2, warps_y=4),
            dict(kernel=1, strip=2, kchunk=5, warps_x=1, warps_y=8),
            dict(kernel=1, strip=4, kchunk=32, warps_x=2, warps_y=4),
            dict(kernel=1, strip=4, kchunk=1, warps_x=8, warps_y=1),
            dict(kernel=1, strip=2, kchunk=1000, warps_x=4, warps_y=2),
            # fused single-sweep step (H and E in one launch, double-buffered state)
            dict(kernel=2, strip=1, kchunk=32, warps_x=2, warps_y=2),
            dict(kernel=2, strip=2, kchunk=2, warps_x=1, warps_y=4),
            dict(kernel=2, strip=2, kchunk=7, warps_x=4, warps_y=1),
            dict(kernel=2, strip=3, kchunk=5, warps_x=2, warps_y=2),
            dict(kernel=2, strip=4, kchunk=1000, warps_x=1, warps_y=2),
            dict(kernel=2, strip=2, kchunk=1, warps_x=2, warps_y=2),
            # the same fused step with TMA-staged operands (cp.async.bulk.tensor + mbarrier ring)
            dict(kernel=3, strip=1, kchunk=32, warps_x=1, warps_y=4, stages=3),
            dict(kernel=3, strip=2, kchunk=5, warps_x=2, warps_y=2, stages=2),
            dict(kernel=3, strip=2, kchunk=1000, warps_x=1, warps_y=8, stages=4),
            dict(kernel=3, strip=1, kchunk=2, warps_x=2, warps_y=4, stages=8),
            # two time steps per sweep (odd step counts finish with one single-step sweep)
            dict(kernel=4),
            dict(kernel=4, kchunk=4, warps_y=8, stages=2),
            dict(kernel=4, kchunk=5, warps_y=16, stages=3),
            dict(kernel=4, kchunk=1000, warps_y=12, stages=4),
            # persistent, warp-specialised cooperative form with flow control between the blocks of a round
            dict(kernel=4, persistent=1, window=1),
            dict(kernel=4, persistent=1, window=8, warps_y=16, stages=2),
            dict(kernel=4, persistent=1, window=2, stages=5),
            # in place on the rolling window of plane slots (what a context does when a second set does not fit)
            dict(kernel=4, rolling=1),
            dict(kernel=4, rolling=1, kchunk=4, warps_y=12, stages=2),
            dict(kernel=4, rolling=1, kchunk=1000, warps_y=16)]
GRIDS = [(0.037, 0.029, 0.023), (0.013, 0.011, 0.009), (0.05, 0.04, 0.03), (0.065, 0.033, 0.012),
         (0.034, 0.066, 0.007),
         # I a multiple of the tile width: the last block in x holds the single column i = I
         (0.0645, 0.0125, 0.0095), (0.1285, 0.0145, 0.0055)]


def configure(ctx, variant):
    for k, v in variant.items():
        ctx.set_option(k, v)


def test_upload_download_roundtrip(F, oracle):
    p = F.make_params(0.037, 0.029, 0.023, 0.001, 6e-13, 1.2e-10, 2, 1)
    q = to_oracle_params(oracle, p)
    f = random_state(oracle, q, 5)
    with F.Context(p) as ctx:
        info = ctx.info()
        assert info["pitch"] % 16 == 0 and info["pitch"] >= p.maxi + 1
        ctx.upload(upper(f))
        assert_fields_equal(ctx.download(), f, "roundtrip")
        assert ctx.checksum() == F.checksum_host(upper(f))


@pytest.mark.parametrize("variant", VARIANTS[:4] + VARIANTS[7:8])
@pytest.mark.parametrize("dims", GRIDS[:3])
def test_operators_match_oracle(F, oracle, dims, variant):
    """update_H_field / update_E_field / set_source one call at a time (main.c:431, :469, :712)."""
    o = oracle.restatement()
    p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, 1)
    q = to_oracle_params(oracle, p)
    f = random_state(oracle, q, 11)
    with F.Context(p) as ctx:
        configure(ctx, variant)
        ctx.upload(upper(f))
        ctx.update_H_field(); o.update_h(q, f)
        assert_fields_equal(ctx.download(), f, "after H")
        ctx.update_E_field(); o.update_e(q, f)
        assert_fields_equal(ctx.download(), f, "after E")
        ctx.set_source(3 * 6e-13); o.set_source(q, f, 3 * 6e-13)
        assert_fields_equal(ctx.download(), f, "after source")
        ctx.update_H_field(); o.update_h(q, f)
        ctx.update_E_field(); o.update_e(q, f)
        assert_fields_equal(ctx.download(), f, "second step")


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("dims", GRIDS)
@pytest.mark.parametrize("mode", [0, 1])
def test_run_matches_oracle_from_random_state(F, oracle, dims, variant, mode):
    """fdtd_run (source and PEC fused into the kernels) == the loop body main.c:770-779."""
    o = oracle.restatement()
    p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, mode)
    q = to_oracle_params(oracle, p)
    f = random_state(oracle, q, 1234)
    with F.Context(p) as ctx:
        configure(ctx, variant)
        ctx.upload(upper(f))
        t_gpu = ctx.run(3, 0.0)
        t_cpu = o.run(q, f, 3)
        assert t_gpu == t_cpu
        assert_fields_equal(ctx.download(), f, "3 steps")
        t_gpu = ctx.run(4, t_gpu)
        t_cpu = o.run(q, f, 4, t_cpu)
        assert t_gpu == t_cpu
        assert_fields_equal(ctx.download(), f, "7 steps")
        assert ctx.checksum() == F.checksum_host(upper(f))


@pytest.mark.parametrize("dims", [(0.001, 0.001, 0.001), (0.002, 0.001, 0.003), (0.016, 0.001, 0.001),
                                  (0.001, 0.017, 0.002), (0.031, 0.032, 0.001), (0.033, 0.031, 0.002)])
def test_degenerate_grids(F, oracle, dims):
    """1-cell-thick cavities and rows shorter/longer than a warp (validation mode: no patch needed)."""
    o = oracle.restatement()
    p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, 0)
    q = to_oracle_params(oracle, p)
    for variant in VARIANTS:
        f = random_state(oracle, q, 77)
        with F.Context(p) as ctx:
            configure(ctx, variant)
            ctx.upload(upper(f))
            ctx.run(5, 0.0)
            o.run(q, f, 5)
            assert_fields_equal(ctx.download(), f, f"{dims} {variant}")


def test_source_patch_must_fit(F):
    p = F.make_params(0.003, 0.003, 0.003, 0.001, 6e-13, 1.2e-10, 2, 1)
    with pytest.raises(F.FdtdError) as e:
        F.Context(p)
    assert e.value.code == -1


@pytest.mark.parametrize("name", ["stock_validation", "stock_computation", "ragged_50x39x29_computation",
                                  "random_37x28x23_computation", "random_33x17x9_validation",
                                  "cube128_computation_200"])
@pytest.mark.parametrize("kernel", [1, 2, 3, 4])
def test_golden_runs(F, golden, tmp_path, name, kernel):
    """Whole runs against digests of the reference's own output (tests/golden/digests.json)."""
    g = golden[name]
    path = tmp_path / "p.txt"
    path.write_text("\n".join(g["params"]))
    p = F.load_parameters(path)
    assert list(p.dims()) == g["grid"]
    with F.Context(p) as ctx:
        ctx.set_option("kernel", kernel)
        if g["init"].startswith("random:"):
            rng = np.random.default_rng(int(g["init"].split(":")[1]))
            f = {n: rng.uniform(-1.0, 1.0, size=s) for n, s in F.field_shapes(p).items()}
            ctx.upload(f)
        elif p.mode == 0:
            ctx.set_initial_conditions()
        t = ctx.run(g["steps"], 0.0)
        assert repr(t) == g["t_end"]
        out = ctx.download()
        for k, want in g["sha256"].items():
            assert digest(out[k[0].upper() + k[1:]]) == want, k
        for v, k in enumerate(F.DUMP_NAMES):
            assert digest(ctx.aggregate(v)) == g["dump_sha256"][k], k


def test_golden_arrays_small_case(F, golden):
    """Same as above with the arrays themselves, so a mismatch can be localised."""
    import os
    from conftest import GOLDEN_DIR
    g = golden["random_37x28x23_computation"]
    want = np.load(os.path.join(GOLDEN_DIR, "random_37x28x23_computation.npz"))
    p = F.make_params(0.037, 0.029, 0.023, 0.001, 6e-13, 1.2e-10, 2, 1)
    rng = np.random.default_rng(1234)
    f = {n: rng.uniform(-1.0, 1.0, size=s) for n, s in F.field_shapes(p).items()}
    with F.Context(p) as ctx:
        ctx.upload(f)
        ctx.run(g["steps"], 0.0)
        assert_fields_equal(ctx.download(), {k: want[k] for k in want.files}, "golden arrays")


@pytest.mark.parametrize("kernel", [1, 2, 3, 4])
def test_config2_cube256_1000_steps(F, golden, tmp_path, kernel):
    """BASELINE.json configs[1]: 256^3, computation mode, 1000 steps, bit-exact vs the reference."""
    g = golden.get("cube256_computation_1000")
    if g is None:
        pytest.skip("fixture not generated (oracle/make_golden.py --full)")
    path = tmp_path / "p.txt"
    path.write_text("\n".join(g["params"]))
    p = F.load_parameters(path)
    assert p.dims() == (256, 256, 256) and F.step_count(p) == 1000 == g["steps"]
    with F.Context(p) as ctx:
        ctx.set_option("kernel", kernel)
        t = ctx.run(1000, 0.0)
        assert repr(t) == g["t_end"]
        out = ctx.download()
        for k, want in g["sha256"].items():
            assert digest(out[k[0].upper() + k[1:]]) == want, k


def test_aggregate_matches_oracle(F, oracle):
    o = oracle.restatement()
    for dims, mode in (((0.037, 0.029, 0.023), 1), ((0.013, 0.034, 0.009), 0)):
        p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, mode)
        q = to_oracle_params(oracle, p)
        f = random_state(oracle, q, 8)
        with F.Context(p) as ctx:
            ctx.upload(upper(f))
            for v in range(6):
                assert bits_equal(ctx.aggregate(v), o.aggregate(q, f, v)), v


@pytest.mark.parametrize("kernel", [1, 2, 3, 4])
@pytest.mark.parametrize("mode", [0, 1])
def test_propagate_matches_reference_dumps(F, golden, tmp_path, mode, kernel):
    """fdtd_propagate == propagate_fields (main.c:755-799): same dump files, same variables in the
    same order, same contents (validation mode adds aEy, aHx, aHz), same final state."""
    g = golden["propagate_tiny"][f"mode{mode}"]
    path = tmp_path / "p.txt"
    path.write_text("\n".join(g["params"]))
    p = F.load_parameters(path)
    log = []
    with F.Context(p) as ctx:
        ctx.set_option("kernel", kernel)
        if mode == 0:
            ctx.set_initial_conditions()
        steps, _ = ctx.propagate(
            on_begin=lambda it, dims, k0: log.append({"file": "r/result%04d.silo" % it, "vars": {}, "order": []}),
            on_variable=lambda name, arr: (log[-1]["vars"].__setitem__(name, digest(arr)),
                                           log[-1]["order"].append(name)))
        assert steps == g["steps"]
        out = ctx.download()
    order = ["ex", "ey", "ez", "hx", "hy", "hz"] + (["aEy", "aHx", "aHz"] if mode == 0 else [])
    assert [d["file"] for d in log] == [d["file"] for d in g["dumps"]]
    for got, want in zip(log, g["dumps"]):
        assert got["order"] == order
        assert got["vars"] == want["vars"], got["file"]
    for k, want in g["final_sha256"].items():
        assert digest(out[k[0].upper() + k[1:]]) == want, k


def test_pattern_fill_matches_host_mirror(F):
    p = F.make_params(0.037, 0.029, 0.023, 0.001, 6e-13, 1.2e-10, 2, 1)
    with F.Context(p) as ctx:
        ctx.fill_test_pattern(2024)
        got = ctx.download()
        want = F.pattern_host(p, 2024)
        for k in want:
            assert bits_equal(got[k], want[k]), k
        assert ctx.checksum() == F.checksum_host(want)


def test_variants_agree_at_larger_size(F):
    """Size-independent property: every kernel variant produces the same state (checksums) from the
    same pattern; 256 x 200 x 120 cells, 6 steps, computation mode."""
    p = F.make_params(0.256, 0.2, 0.12, 0.001, 6e-13, 1.2e-10, 2, 1)
    sums = []
    for variant in VARIANTS:
        with F.Context(p) as ctx:
            configure(ctx, variant)
            ctx.fill_test_pattern(9)
            ctx.run(6, 0.0)
            sums.append(ctx.checksum())
    assert all(s == sums[0] for s in sums[1:])


def test_linearity_power_of_two_scaling(F, oracle):
    """The update is linear: doubling the state doubles the result exactly (validation mode, no
    source) -- a property that holds at any size."""
    p = F.make_params(0.064, 0.05, 0.033, 0.001, 6e-13, 1.2e-10, 2, 0)
    q = to_oracle_params(oracle, p)
    f = random_state(oracle, q, 3)
    with F.Context(p) as a, F.Context(p) as b:
        a.upload(upper(f))
        b.upload(upper({k: 2.0 * v for k, v in f.items()}))
        a.run(25, 0.0)
        b.run(25, 0.0)
        fa, fb = a.download(), b.download()
        for k in fa:
            assert bits_equal(2.0 * fa[k], fb[k]), k


def test_random_shapes_and_launch_configs(F, oracle):
    """40 seeded random cavities (1..70 cells per axis, both modes where the source patch fits) x random
    kernel / strip / chunk / tile / ring settings, 1-6 steps from a random state: all bit-exact."""
    o = oracle.restatement()
    rng = np.random.default_rng(20261018)
    done = 0
    while done < 40:
        nx, ny, nz = (int(v) for v in rng.integers(1, 71, size=3))
        mode = int(rng.integers(0, 2))
        dims = ((nx + .5) * 1e-3, (ny + .5) * 1e-3, (nz + .5) * 1e-3)
        p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, mode)
        assert p.dims() == (nx, ny, nz)
        if mode == 1:
            plan = F.source_plan(p)
            if plan.i0 < 0 or plan.j0 < 0 or plan.i1 > nx or plan.j1 > ny:
                continue
        kernel = int(rng.integers(0, 4))
        variant = dict(kernel=kernel, kchunk=int(rng.choice([1, 2, 3, 7, 32, 1000])))
        if kernel == 1:
            wx, wy = [(1, 8), (2, 4), (4, 2), (8, 1), (2, 2), (1, 1)][int(rng.integers(0, 6))]
            variant.update(strip=int(rng.choice([1, 2, 4])), warps_x=wx, warps_y=wy)
        elif kernel == 2:
            wx, wy = [(1, 4), (2, 2), (4, 1), (1, 2), (1, 1)][int(rng.integers(0, 5))]
            variant.update(strip=int(rng.integers(1, 5)), warps_x=wx, warps_y=wy, prefetch=int(rng.integers(0, 5)))
        elif kernel == 3:
            wx, wy = [(1, 8), (2, 4), (4, 2), (2, 2), (1, 4), (4, 4), (3, 2), (1, 1)][int(rng.integers(0, 8))]
            variant.update(strip=2 if wx * wy > 8 else int(rng.integers(1, 3)), warps_x=wx, warps_y=wy,
                           stages=int(rng.integers(2, 7)))
        steps = int(rng.integers(1, 7))
        q = to_oracle_params(oracle, p)
        f = random_state(oracle, q, int(rng.integers(0, 1 << 30)))
        with F.Context(p) as ctx:
            configure(ctx, variant)
            ctx.upload(upper(f))
            t_gpu = ctx.run(steps, 0.0)
            t_cpu = o.run(q, f, steps)
            assert t_gpu == t_cpu
            assert_fields_equal(ctx.download(), f, f"{(nx, ny, nz)} mode {mode} {variant} {steps} steps")
        done += 1


def test_two_step_kernel_random_shapes(F, oracle):
    """kernel 4 (two time steps per sweep): 40 seeded random cavities x chunk length / block height /
    ring depth, 1-9 steps from a random state, both modes: bit-exact against the oracle."""
    o = oracle.restatement()
    rng = np.random.default_rng(4242)
    done = 0
    while done < 40:
        nx, ny, nz = (int(v) for v in rng.integers(1, 90, size=3))
        mode = int(rng.integers(0, 2))
        dims = ((nx + .5) * 1e-3, (ny + .5) * 1e-3, (nz + .5) * 1e-3)
        p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, mode)
        if mode == 1:
            plan = F.source_plan(p)
            if plan.i0 < 0 or plan.j0 < 0 or plan.i1 > nx or plan.j1 > ny:
                continue
        variant = dict(kernel=4, kchunk=int(rng.choice([1, 2, 3, 4, 7, 32, 64, 1000])), warps_y=int(rng.choice([8, 12, 16])),
                       stages=int(rng.integers(2, 6)))
        steps = int(rng.integers(1, 10))
        q = to_oracle_params(oracle, p)
        f = random_state(oracle, q, int(rng.integers(0, 1 << 30)))
        with F.Context(p) as ctx:
            configure(ctx, variant)
            ctx.upload(upper(f))
            t_gpu = ctx.run(steps, 0.0)
            t_cpu = o.run(q, f, steps)
            assert t_gpu == t_cpu
            assert_fields_equal(ctx.download(), f, f"{(nx, ny, nz)} mode {mode} {variant} {steps} steps")
        done += 1
