import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "digests.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def oracle():
    """The CPU checkers (test infrastructure): oracle.restatement(), oracle.reference()."""
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def F():
    """The product's ctypes view of libfdtd_b200.so; builds the library when it is missing."""
    lib = os.path.join(ROOT, "fdtd-maxwell-microwave-oven_b200", "libfdtd_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    import fdtd_b200
    return fdtd_b200


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def to_oracle_params(O, p):
    """fdtd_params -> oracle_params (same numbers; the grid is re-derived by the oracle)."""
    q = O.make_params(p.length, p.width, p.height, p.spatial_step, p.time_step, p.simulation_time,
                      p.sampling_rate, p.mode)
    assert q.dims() == p.dims()
    return q


def upper(fields):
    return {k[0].upper() + k[1:]: v for k, v in fields.items()}


def lower(fields):
    return {k.lower(): v for k, v in fields.items()}
