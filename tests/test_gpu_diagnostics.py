"""Diagnostics next to the path: energy and the analytic-mode error.
They are parallel reductions, so the tolerance is rounding (1e-12 relative), not zero."""
import numpy as np
import pytest

from conftest import to_oracle_params, upper

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.mark.parametrize("dims", [(0.037, 0.029, 0.023), (0.05, 0.04, 0.03)])
def test_energy_matches_oracle(F, oracle, dims):
    o = oracle.restatement()
    p = F.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, 1)
    q = to_oracle_params(oracle, p)
    f = oracle.alloc_fields(*q.dims(), rng=np.random.default_rng(21))
    with F.Context(p) as ctx:
        ctx.upload(upper(f))
        ctx.run(3, 0.0); o.run(q, f, 3)
        e_ref, h_ref = o.energy(q, f)                    # as coded (main.c:627 slip included)
        e_gpu, h_gpu = ctx.energy(as_coded=True)
        assert abs(e_gpu - e_ref) <= RTOL * abs(e_ref)
        assert abs(h_gpu - h_ref) <= RTOL * abs(h_ref)
        # intended zone average of Ez: differs from the as-coded value on random fields, same H part
        e_int, h_int = ctx.energy(as_coded=False)
        assert abs(h_int - h_gpu) <= RTOL * abs(h_gpu) and abs(e_int - e_gpu) > 1e-6 * abs(e_gpu)
        nx, ny, nz = q.dims()
        dv = q.spatial_step ** 3
        ez = f["ez"]
        mez = (ez[:, :-1, :-1] + ez[:, 1:, :-1] + ez[:, :-1, 1:] + ez[:, 1:, 1:]) / 4.
        ex, ey = f["ex"], f["ey"]
        mex = (ex[:-1, :-1, :] + ex[1:, :-1, :] + ex[:-1, 1:, :] + ex[1:, 1:, :]) / 4.
        mey = (ey[:-1, :, :-1] + ey[:-1, :, 1:] + ey[1:, :, :-1] + ey[1:, :, 1:]) / 4.
        want = ((mex ** 2).sum() + (mey ** 2).sum() + (mez ** 2).sum()) * dv * 8.854e-12 / 2.
        assert abs(e_int - want) <= 1e-11 * abs(want)


def test_validation_error_matches_oracle_and_physics(F, oracle):
    """e_r of the TE101 mode (description.pdf eq. 2) from the device equals the one computed from the
    oracle's validation fields; and the mode really is a solution: the error stays small."""
    o = oracle.restatement()
    p = F.make_params(0.05, 0.05, 0.05, 0.001, 6e-13, 1.2e-10, 2, 0)
    q = to_oracle_params(oracle, p)
    f = oracle.alloc_fields(*q.dims())
    o.set_initial_conditions(q, f)
    with F.Context(p) as ctx:
        ctx.set_initial_conditions()
        t_gpu = ctx.run(200, 0.0)
        t_cpu = o.run(q, f, 200)
        assert t_gpu == t_cpu
        t_eval = t_cpu - q.time_step       # the reference evaluates at the step's own time (main.c:783)
        sums, rel = ctx.validation_error(t_eval)
        v = o.validation_fields(q, f, t_eval)
        for n, key in enumerate(("ey", "hx", "hz")):
            num = float((v[key] ** 2).sum())
            den = float(((v[key] + f[key]) ** 2).sum())
            assert abs(sums[2 * n] - num) <= 1e-10 * num
            assert abs(sums[2 * n + 1] - den) <= 1e-10 * den
            assert abs(rel[n] - np.sqrt(num / den)) <= 1e-10
        assert rel[0] < 0.05          # a 50-cell cavity after 200 steps: a few percent at most
