"""The oracle itself: restatement (oracle/fdtd_oracle.c) against the golden vectors generated from
the compiled reference, and -- where oracle/_ref exists -- against the reference directly."""
import hashlib
import os

import numpy as np
import pytest

from conftest import bits_equal


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def initial_state(O, chk, p, spec):
    if spec.startswith("random:"):
        return O.alloc_fields(*p.dims(), rng=np.random.default_rng(int(spec.split(":")[1])))
    f = O.alloc_fields(*p.dims())
    if p.mode == 0:
        chk.set_initial_conditions(p, f)
    return f


FAST_CASES = ["stock_validation", "stock_computation", "ragged_50x39x29_computation",
              "random_37x28x23_computation", "random_33x17x9_validation"]


@pytest.mark.parametrize("name", FAST_CASES)
def test_restatement_matches_golden(oracle, golden, tmp_path, name):
    g = golden[name]
    o = oracle.restatement()
    p = o.load_parameters(oracle.write_params(tmp_path / "p.txt", g["params"]))
    assert list(p.dims()) == g["grid"]
    f = initial_state(oracle, o, p, g["init"])
    t_end = o.run(p, f, g["steps"])
    assert repr(t_end) == g["t_end"]
    for k, v in f.items():
        assert digest(v) == g["sha256"][k], k
    for v, k in enumerate(oracle.FIELD_NAMES):
        assert digest(o.aggregate(p, f, v)) == g["dump_sha256"][k], k
    if p.mode == 1:
        assert list(oracle.source_bounds(p)) == g["source_bounds"]
        assert repr(oracle.source_zte(p)) == g["z_te"]


def test_step_counts_follow_the_float_bound(oracle, golden, tmp_path):
    o = oracle.restatement()
    for name in FAST_CASES[:3] + ["cube128_computation_200"]:
        p = o.load_parameters(oracle.write_params(tmp_path / "p.txt", golden[name]["params"]))
        assert oracle.step_count(p) == golden[name]["steps"]
    # 6e-10 (as float) admits exactly 1000 steps of 6e-13, 5.99e-10 -> 999
    for text, want in (("0.0000000006", 1000), ("0.000000000599", 999)):
        p = o.load_parameters(oracle.write_params(tmp_path / "p.txt",
                              ("0.01", "0.01", "0.01", "0.001", "0.0000000000006", text, "1", "1")))
        assert oracle.step_count(p) == want


def test_float_parameters_decide_the_grid(oracle, tmp_path):
    o = oracle.restatement()
    # SURVEY.md Appendix B.5 / B.10
    for text, cells in (("0.05", 50), ("0.04", 39), ("0.03", 29), ("0.256", 256), ("1.024", 1024),
                        ("2.048", 2048), ("0.1", 100)):
        p = o.load_parameters(oracle.write_params(tmp_path / "p.txt",
                              (text, text, text, "0.001", "6e-13", "1e-10", "1", "1")))
        assert p.dims() == (cells,) * 3, text


def test_known_answers_from_the_report(oracle):
    # description.pdf p.4: Z_TE = 532.7884 Ohm for the 0.05 m stock guide; patch 7x7 at 1 mm
    p = oracle.make_params(0.05, 0.05, 0.05, 0.001, 6e-13, 1.2e-10, 2, 1)
    assert abs(oracle.source_zte(p) - 532.7884) < 1e-4
    assert oracle.source_bounds(p) == (21, 28, 21, 28)
    for n, lo in ((0.256, 124), (1.024, 508), (2.048, 1020)):
        q = oracle.make_params(n, n, n, 0.001, 6e-13, 1.2e-10, 2, 1)
        assert oracle.source_bounds(q) == (lo, lo + 7, lo, lo + 7)


def test_validation_mode_leaves_three_components_zero(oracle):
    o = oracle.restatement()
    p = oracle.make_params(0.05, 0.05, 0.05, 0.001, 6e-13, 1.2e-10, 2, 0)
    f = oracle.alloc_fields(*p.dims())
    o.set_initial_conditions(p, f)
    o.run(p, f, 20)
    for k in ("ex", "ez", "hy"):
        assert not f[k].any()
    # the k = K and i = I faces of the initial Ey are tiny but NOT zero
    assert f["ey"][-1].any() and abs(f["ey"][-1]).max() < 1e-6


needs_ref = pytest.mark.skipif(not os.path.exists("/root/reference/main.c") and
                               not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle",
                                                               "_ref", "libfdtd_ref.so")),
                               reason="oracle/_ref was not built (reference not mounted)")


@needs_ref
@pytest.mark.parametrize("dims,mode,seed", [((0.013, 0.011, 0.009), 1, 1), ((0.021, 0.016, 0.012), 1, 2),
                                            ((0.009, 0.013, 0.017), 0, 3), ((0.001, 0.001, 0.001), 0, 4)])
def test_restatement_matches_reference_functions(oracle, dims, mode, seed):
    o, r = oracle.restatement(), oracle.reference()
    assert r is not None
    p = oracle.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, mode)
    a = oracle.alloc_fields(*p.dims(), rng=np.random.default_rng(seed))
    b = {k: v.copy() for k, v in a.items()}
    for step in range(3):
        t = step * 6e-13
        if mode == 1:
            o.set_source(p, a, t); r.set_source(p, b, t)
        o.update_h(p, a); r.update_h(p, b)
        if mode == 1:
            o.set_source(p, a, t); r.set_source(p, b, t)
        o.update_e(p, a); r.update_e(p, b)
        for k in a:
            assert bits_equal(a[k], b[k]), (step, k)
    for v in range(6):
        assert bits_equal(o.aggregate(p, a, v), r.aggregate(p, b, v))
    va, vb = o.validation_fields(p, a, 1.3e-11), r.validation_fields(p, b, 1.3e-11)
    for k in va:
        assert bits_equal(va[k], vb[k]), k
    ia, ib = oracle.alloc_fields(*p.dims()), oracle.alloc_fields(*p.dims())
    o.set_initial_conditions(p, ia); r.set_initial_conditions(p, ib)
    assert bits_equal(ia["ey"], ib["ey"])


@needs_ref
def test_reference_loader_agrees(oracle, tmp_path):
    o, r = oracle.restatement(), oracle.reference()
    for nums in (oracle.STOCK_PARAMS, ("0.04", "0.03", "0.05", "0.0005", "3e-13", "2e-11", "7", "1")):
        path = oracle.write_params(tmp_path / "p.txt", nums)
        a, b = o.load_parameters(path), r.load_parameters(path)
        for name, _ in oracle.Params._fields_:
            assert getattr(a, name) == getattr(b, name), name


@needs_ref
def test_reference_propagate_matches_golden(oracle, golden, tmp_path):
    """The whole reference program loop (dump cadence and contents) is what the fixture says."""
    r = oracle.reference()
    for mode in (0, 1):
        g = golden["propagate_tiny"][f"mode{mode}"]
        p = r.load_parameters(oracle.write_params(tmp_path / "p.txt", g["params"]))
        f = initial_state(oracle, r, p, "params")
        log = []

        def on_var(kind, name, arr):
            if kind == 0:
                log.append({"file": name, "vars": {}})
            elif kind == 1:
                log[-1]["vars"][name] = digest(arr)
        r.propagate(p, f, on_var)
        assert log == g["dumps"]
        assert {k: digest(v) for k, v in f.items()} == g["final_sha256"]


@needs_ref
def test_energy_restatement_matches_reference(oracle):
    """calculate_E_energy / calculate_H_energy (main.c:602-668) incl. the Ez[kHz] slip at main.c:627."""
    o, r = oracle.restatement(), oracle.reference()
    for dims in ((0.013, 0.011, 0.009), (0.021, 0.016, 0.012)):
        p = oracle.make_params(*dims, 0.001, 6e-13, 1.2e-10, 2, 1)
        f = oracle.alloc_fields(*p.dims(), rng=np.random.default_rng(3))
        assert o.energy(p, f) == r.energy(p, f)
