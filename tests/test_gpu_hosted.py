"""fdtd_run_hosted: host arrays in, K steps, host arrays out.  The wavefront over z-chunks (upload,
stepping and download overlapped) must leave exactly what upload + fdtd_run + download leaves, which
in turn is the oracle's result, bit for bit -- for every chunk size, including chunks of one block's
prologue only, chunk counts that do not divide the planes, more steps than chunks and long runs that
ramp in and out around whole-grid steps."""
import numpy as np
import pytest

from conftest import bits_equal, upper

pytestmark = pytest.mark.gpu


def _case(F, oracle, dims, mode, steps, opts, t0=0.0):
    o = oracle.restatement()
    args = tuple((d + .5) * 1e-3 for d in dims) + (0.001, 6e-13, 1e-9, 1 << 20, mode)
    p, q = F.make_params(*args), oracle.make_params(*args)
    assert p.dims() == dims
    f = oracle.alloc_fields(*dims, rng=np.random.default_rng(21))
    host = F.PinnedArrays(p)
    for k, v in upper(f).items():
        host.arrays[k][...] = v
    with F.Context(p, device=0) as ctx:
        for k, v in opts.items():
            ctx.set_option(k, v)
        t_gpu = ctx.run_hosted(host.arrays, steps, t0)
        # the context holds the same final state as the host arrays
        dev = ctx.download()
    t_cpu = o.run(q, f, steps, t0)
    assert t_gpu == t_cpu
    for k, want in f.items():
        name = k[0].upper() + k[1:]
        assert bits_equal(host.arrays[name], want), (dims, steps, opts, k, "host arrays")
        assert bits_equal(dev[name], want), (dims, steps, opts, k, "device state")
    host.close()


@pytest.mark.parametrize("chunk", [2, 3, 5, 8, 64])
@pytest.mark.parametrize("steps", [1, 2, 7, 20])
def test_wavefront_matches_oracle(F, oracle, chunk, steps):
    _case(F, oracle, (37, 28, 23), 1, steps, dict(host_chunk=chunk))


@pytest.mark.parametrize("opts", [dict(kernel=2, strip=2, kchunk=8, warps_x=2, warps_y=2, host_chunk=4),
                                  dict(kernel=3, strip=2, kchunk=4, warps_x=2, warps_y=2, stages=3, host_chunk=4),
                                  dict(host_chunk=0),
                                  dict(host_pipeline=0),
                                  dict(kernel=1), dict(kernel=0)])
@pytest.mark.parametrize("mode", [0, 1])
def test_kernels_and_modes(F, oracle, opts, mode):
    _case(F, oracle, (64, 21, 19), mode, 9, opts, t0=3 * 6e-13)


def test_long_run_ramps_in_and_out(F, oracle):
    """more than 64 steps: wavefront for the first and last 32, whole-grid steps in between"""
    _case(F, oracle, (33, 30, 12), 1, 101, dict(host_chunk=3))


def test_degenerate_grids(F, oracle):
    _case(F, oracle, (40, 33, 1), 1, 5, dict(host_chunk=2))
    _case(F, oracle, (1, 1, 9), 0, 4, dict(host_chunk=2))


def test_repeated_calls_continue_the_run(F, oracle):
    o = oracle.restatement()
    dims = (37, 28, 23)
    args = tuple((d + .5) * 1e-3 for d in dims) + (0.001, 6e-13, 1e-9, 1 << 20, 1)
    p, q = F.make_params(*args), oracle.make_params(*args)
    f = oracle.alloc_fields(*dims, rng=np.random.default_rng(22))
    host = F.PinnedArrays(p)
    for k, v in upper(f).items():
        host.arrays[k][...] = v
    with F.Context(p, device=0) as ctx:
        ctx.set_option("host_chunk", 4)
        t = ctx.run_hosted(host.arrays, 5, 0.0)
        t = ctx.run_hosted(host.arrays, 4, t)
        t = ctx.run(3, t)                     # the device state is current too
        got = ctx.download()
    o.run(q, f, 12)
    for k, want in f.items():
        assert bits_equal(got[k[0].upper() + k[1:]], want), k
    host.close()


def test_full_size_checksum(F):
    """512^3, 20 steps: the wavefront and the plain sequence agree by checksum (no host oracle at this size)"""
    n = 512
    p = F.make_params(n * 1e-3, n * 1e-3, n * 1e-3, 0.001, 6e-13, 1e-9, 1 << 30, 1)
    assert p.dims() == (n, n, n)
    host = F.PinnedArrays(p)
    sums = []
    with F.Context(p, device=0) as ctx:
        for pipeline in (1, 0):
            ctx.set_option("host_pipeline", pipeline)
            ctx.fill_test_pattern(99)
            ctx.download_slab(host.arrays)
            ctx.run_hosted(host.arrays, 20, 0.0)
            ctx.fill_test_pattern(1)          # forget the device state: the host arrays are the result
            ctx.upload_slab(host.arrays)
            sums.append(ctx.checksum())
    host.close()
    assert sums[0] == sums[1]
