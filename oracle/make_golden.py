#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the COMPILED REFERENCE (oracle/_ref).

Run in the build container, where /root/reference is mounted:

    python oracle/make_golden.py            # everything but the long case (~1 min)
    python oracle/make_golden.py --full     # also 256^3 x 1000 steps (BASELINE.json configs[1], ~8 min)

The reference has no golden vectors of its own, so these are outputs of the
reference itself, run here; the GPU box has no /root/reference and only reads the fixtures.
Digest = sha256 over the raw little-endian float64 bytes of each dense array.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (params.txt numbers, initial state, steps or None for the params' own count)
CASES = {
    "stock_validation": (O.STOCK_PARAMS, "params", None),
    "stock_computation": (O.STOCK_PARAMS[:7] + ("1",), "params", None),
    "ragged_50x39x29_computation": (("0.05", "0.04", "0.03", "0.001", "0.0000000000006",
                                     "0.00000000012", "2", "1"), "params", None),
    "random_37x28x23_computation": (("0.037", "0.029", "0.023", "0.001", "0.0000000000006",
                                     "0.00000000012", "2", "1"), "random:1234", 10),
    "random_33x17x9_validation": (("0.033", "0.017", "0.009", "0.001", "0.0000000000006",
                                   "0.00000000012", "2", "0"), "random:99", 7),
    "cube128_computation_200": (("0.128", "0.128", "0.128", "0.001", "0.0000000000006",
                                 "0.00000000012", "50", "1"), "params", None),
}
FULL_CASES = {
    # BASELINE.json configs[1]: 256^3, computation mode, exactly 1000 steps
    "cube256_computation_1000": (("0.256", "0.256", "0.256", "0.001", "0.0000000000006",
                                  "0.0000000006", "1000000", "1"), "params", None),
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def initial_state(chk, p, spec):
    if spec.startswith("random:"):
        return O.alloc_fields(*p.dims(), rng=np.random.default_rng(int(spec.split(":")[1])))
    f = O.alloc_fields(*p.dims())
    if p.mode == 0:
        chk.set_initial_conditions(p, f)
    return f


def run_case(ref, name, numbers, init, steps):
    path = O.write_params(os.path.join("/tmp", name + ".txt"), numbers)
    p = ref.load_parameters(path)
    f = initial_state(ref, p, init)
    n = O.step_count(p) if steps is None else steps
    t0 = time.time()
    t_end = ref.run(p, f, n)
    rec = {"params": list(numbers), "grid": list(p.dims()), "init": init, "steps": n,
           "t_end": repr(t_end), "sha256": {k: digest(v) for k, v in f.items()},
           "l2": {k: repr(float(np.sqrt(np.sum(v * v)))) for k, v in f.items()}}
    if p.mode == 1:
        rec["source_bounds"] = list(O.source_bounds(p))
        rec["z_te"] = repr(O.source_zte(p))
    rec["dump_sha256"] = {O.FIELD_NAMES[v]: digest(ref.aggregate(p, f, v)) for v in range(6)}
    print(f"{name}: grid {p.dims()} steps {n} ({time.time() - t0:.1f}s)")
    return rec, p, f


def propagate_case(ref):
    """Whole reference program loop (propagate_fields) on a tiny grid: which dumps, what is in them."""
    out = {}
    for mode in (0, 1):
        nums = ("0.012", "0.011", "0.010", "0.001", "0.0000000000006", "0.000000000012", "3", str(mode))
        p = ref.load_parameters(O.write_params("/tmp/prop.txt", nums))
        f = initial_state(ref, p, "params")
        log = []

        def on_var(kind, name, arr, log=log):
            if kind == 0:
                log.append({"file": name, "vars": {}})
            elif kind == 1:
                log[-1]["vars"][name] = digest(arr)
        ref.propagate(p, f, on_var)
        out[f"mode{mode}"] = {"params": list(nums), "grid": list(p.dims()),
                              "steps": O.step_count(p), "dumps": log,
                              "final_sha256": {k: digest(v) for k, v in f.items()}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    ref = O.reference()
    if ref is None:
        sys.exit("oracle/_ref/libfdtd_ref.so missing: run `make -C oracle` where /root/reference is mounted")
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "digests.json")
    gold = json.load(open(path)) if os.path.exists(path) else {}
    gold["_about"] = ("sha256 of the reference's own outputs (oracle/_ref = unmodified main.c, "
                      "gcc -std=c99 -O3, glibc %s); regenerate with oracle/make_golden.py"
                      % os.confstr("CS_GNU_LIBC_VERSION"))
    cases = dict(CASES)
    if args.full:
        cases.update(FULL_CASES)
    for name, (numbers, init, steps) in cases.items():
        rec, p, f = run_case(ref, name, numbers, init, steps)
        gold[name] = rec
        if name == "random_37x28x23_computation":
            # one small case with the arrays themselves, so a GPU-box failure can be localised
            np.savez_compressed(os.path.join(GOLD, "random_37x28x23_computation.npz"), **f)
    gold["propagate_tiny"] = propagate_case(ref)
    with open(path, "w") as fh:
        json.dump(gold, fh, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
