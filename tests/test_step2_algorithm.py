"""The two-steps-per-sweep schedule of the default kernel (csrc/fdtd_step2_tma.cuh), restated in numpy
(tests/step2_numpy.py), against the CPU oracle: overlapped tiles with an unusable rim (NaN there must
never reach a stored value), chunks with a two-plane run-in, PEC walls and the source at both time
levels, and slabs that only see two halo planes each way."""
import numpy as np
import pytest

from step2_numpy import step2_sweep

DT = 6e-13


def _setup(O, F, dims, mode):
    args = tuple((d + .5) * 1e-3 for d in dims) + (0.001, DT, 1e-9, 1 << 20, mode)
    q, p = O.make_params(*args), F.make_params(*args)
    assert q.dims() == dims == p.dims()
    ch = DT / (1.25663706143591729538505735331180115367886775975E-6 * 0.001)
    ce = DT / (8.854E-12 * 0.001)
    plan = F.source_plan(p) if mode == 1 else None
    return q, p, ch, ce, plan


def _src(F, p, plan, t):
    if plan is None:
        return None
    ez, hx = F.source_values(p, plan, t)
    return (plan.i0, plan.i1, plan.j0, plan.j1, ez, hx)


@pytest.mark.parametrize("dims,mode,wy,kchunk", [((23, 19, 9), 1, 8, 4), ((40, 30, 7), 1, 8, 3), ((31, 14, 5), 0, 8, 2),
                                                  ((12, 33, 6), 0, 12, 100), ((1, 1, 3), 0, 8, 4), ((57, 9, 4), 1, 16, 1)])
def test_two_step_schedule_equals_reference_loop(F, oracle, dims, mode, wy, kchunk):
    o = oracle.restatement()
    q, p, ch, ce, plan = _setup(oracle, F, dims, mode)
    if plan is not None and (plan.i0 < 0 or plan.j0 < 0 or plan.i1 > dims[0] or plan.j1 > dims[1]):
        pytest.skip("source patch does not fit")
    want = oracle.alloc_fields(*dims, rng=np.random.default_rng(8))
    a = {k: v.copy() for k, v in want.items()}
    t = 0.0
    for sweep in range(2):                       # four steps = two sweeps
        b = {k: np.full_like(v, np.nan) for k, v in a.items()}
        step2_sweep(a, b, dims, ch, ce, _src(F, p, plan, t), _src(F, p, plan, t + DT), wy=wy, kchunk=kchunk)
        a = b
        t += DT
        t += DT
    o.run(q, want, 4)
    for k, w in want.items():
        assert np.array_equal(a[k].view(np.uint64), w.view(np.uint64)), k


@pytest.mark.parametrize("world", [2, 3])
def test_two_step_slabs_need_two_halo_planes_each_way(F, oracle, world):
    """each slab sweeps its own planes of whole-cavity arrays in which everything beyond two planes below and
    above the slab is NaN; the slabs' planes put together equal the oracle's two steps"""
    o = oracle.restatement()
    dims, mode = (23, 19, 11), 1
    q, p, ch, ce, plan = _setup(oracle, F, dims, mode)
    want = oracle.alloc_fields(*dims, rng=np.random.default_rng(9))
    init = {k: v.copy() for k, v in want.items()}
    o.run(q, want, 2)
    nz = dims[2]
    for rank in range(world):
        k0, k1 = F.slab_range(nz, rank, world)
        a = {k: v.copy() for k, v in init.items()}
        for name, arr in a.items():
            arr[:max(k0 - 2, 0)] = np.nan
            arr[min(k1 + 2, arr.shape[0]):] = np.nan
        b = {k: np.full_like(v, np.nan) for k, v in a.items()}
        step2_sweep(a, b, dims, ch, ce, _src(F, p, plan, 0.0), _src(F, p, plan, DT), kchunk=3, klo=k0, khi=k1)
        top = 1 if rank == world - 1 else 0
        for name in ("ez", "hx", "hy"):
            assert np.array_equal(b[name][k0:k1].view(np.uint64), want[name][k0:k1].view(np.uint64)), (rank, name)
        for name in ("ex", "ey", "hz"):
            assert np.array_equal(b[name][k0:k1 + top].view(np.uint64), want[name][k0:k1 + top].view(np.uint64)), (rank, name)
