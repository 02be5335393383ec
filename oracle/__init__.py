"""ctypes access to the CPU checkers (TEST INFRASTRUCTURE ONLY).

Two libraries with the same call surface:

* ``restatement()`` -- oracle/libfdtd_oracle.so, the C restatement (oracle/fdtd_oracle.c);
* ``reference()``   -- oracle/_ref/libfdtd_ref.so, the unmodified /root/reference/main.c
  compiled by oracle/Makefile (present in the build container and shipped prebuilt to the
  GPU box; ``None`` when it was never built).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIELD_NAMES = ("ex", "ey", "ez", "hx", "hy", "hz")


class Params(C.Structure):
    """oracle_params (oracle/fdtd_oracle.h) == the reference's Parameters, main.c:57-71."""
    _fields_ = [("length", C.c_float), ("width", C.c_float), ("height", C.c_float),
                ("spatial_step", C.c_double), ("time_step", C.c_double),
                ("simulation_time", C.c_float), ("sampling_rate", C.c_uint), ("mode", C.c_int),
                ("nx", C.c_size_t), ("ny", C.c_size_t), ("nz", C.c_size_t)]

    def dims(self):
        return int(self.nx), int(self.ny), int(self.nz)


class FieldPtrs(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in FIELD_NAMES]


def field_shapes(nx, ny, nz):
    """(planes, rows, row length) of the six dense arrays, x fastest (main.c:379-407)."""
    return {"ex": (nz + 1, ny + 1, nx), "ey": (nz + 1, ny, nx + 1), "ez": (nz, ny + 1, nx + 1),
            "hx": (nz, ny, nx + 1), "hy": (nz, ny + 1, nx), "hz": (nz + 1, ny, nx)}


def alloc_fields(nx, ny, nz, rng=None):
    """Six C-contiguous float64 arrays; zeros, or uniform(-1, 1) when a Generator is given."""
    out = {}
    for name, shp in field_shapes(nx, ny, nz).items():
        out[name] = (np.zeros(shp) if rng is None else rng.uniform(-1.0, 1.0, size=shp))
    return out


def _ptrs(fields):
    fp = FieldPtrs()
    for n in FIELD_NAMES:
        a = fields[n]
        assert a.dtype == np.float64 and a.flags.c_contiguous
        setattr(fp, n, a.ctypes.data_as(C.POINTER(C.c_double)))
    return fp


RECORDER = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.POINTER(C.c_double), C.c_size_t)


class Checker:
    """Uniform wrapper over either library (function prefix 'oracle_' or 'ref_')."""

    def __init__(self, lib, prefix, kind):
        self.lib, self.prefix, self.kind = lib, prefix, kind
        P, F, D = C.POINTER(Params), C.POINTER(FieldPtrs), C.POINTER(C.c_double)
        sig = {"load_parameters": (C.c_int, [C.c_char_p, P]),
               "set_initial_conditions": (None, [P, D]),
               "update_h": (None, [P, F]), "update_e": (None, [P, F]),
               "set_source": (None, [P, F, C.c_double]),
               "run": (None, [P, F, C.c_size_t, D]),
               "aggregate": (None, [P, F, C.c_int, D]),
               "validation_fields": (None, [P, F, D, D, D, C.c_double]),
               "energy": (None, [P, F, D])}
        if prefix == "oracle_":
            sig.update({"make_params": (None, [C.c_float, C.c_float, C.c_float, C.c_double, C.c_double,
                                               C.c_float, C.c_uint, C.c_int, P]),
                        "step_count": (C.c_size_t, [P]),
                        "source_bounds": (None, [P, C.POINTER(C.c_long)]),
                        "source_zte": (C.c_double, [P]),
                        "field_sizes": (None, [P, C.POINTER(C.c_size_t)])})
        else:
            sig["propagate"] = (None, [P, F, RECORDER])
        for name, (res, args) in sig.items():
            fn = getattr(lib, prefix + name)
            fn.restype, fn.argtypes = res, args
            setattr(self, "_" + name, fn)

    # -- parameters -----------------------------------------------------------------
    def load_parameters(self, path):
        p = Params()
        if self._load_parameters(os.fsencode(path), C.byref(p)) != 0:
            raise FileNotFoundError(path)
        return p

    # -- operators ------------------------------------------------------------------
    def set_initial_conditions(self, p, fields):
        self._set_initial_conditions(C.byref(p), fields["ey"].ctypes.data_as(C.POINTER(C.c_double)))

    def update_h(self, p, fields):
        self._update_h(C.byref(p), C.byref(_ptrs(fields)))

    def update_e(self, p, fields):
        self._update_e(C.byref(p), C.byref(_ptrs(fields)))

    def set_source(self, p, fields, t):
        self._set_source(C.byref(p), C.byref(_ptrs(fields)), float(t))

    def run(self, p, fields, steps, t0=0.0):
        t = C.c_double(t0)
        self._run(C.byref(p), C.byref(_ptrs(fields)), int(steps), C.byref(t))
        return t.value

    def aggregate(self, p, fields, var):
        nx, ny, nz = p.dims()
        out = np.empty((nz, ny, nx))
        self._aggregate(C.byref(p), C.byref(_ptrs(fields)), int(var),
                        out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def validation_fields(self, p, fields, t):
        shp = field_shapes(*p.dims())
        v = {n: np.empty(shp[n]) for n in ("ey", "hx", "hz")}
        D = C.POINTER(C.c_double)
        self._validation_fields(C.byref(p), C.byref(_ptrs(fields)), v["ey"].ctypes.data_as(D),
                                v["hx"].ctypes.data_as(D), v["hz"].ctypes.data_as(D), float(t))
        return v

    def energy(self, p, fields):
        """(electric, magnetic) energy as coded at main.c:602-668"""
        out = (C.c_double * 2)()
        self._energy(C.byref(p), C.byref(_ptrs(fields)), out)
        return out[0], out[1]

    def propagate(self, p, fields, on_var):
        """reference only: run propagate_fields (main.c:755-799); on_var(kind, name, array|None)."""
        def _cb(kind, name, data, n):
            arr = None
            if kind == 1:
                arr = np.ctypeslib.as_array(data, shape=(n,)).copy()
            on_var(kind, name.decode() if name else None, arr)
        cb = RECORDER(_cb)
        self._propagate(C.byref(p), C.byref(_ptrs(fields)), cb)


_cache = {}


def build(force=False):
    """Compile both checkers (make -C oracle).  Building the checker is not using it."""
    so = os.path.join(HERE, "libfdtd_oracle.so")
    ref = os.path.join(HERE, "_ref", "libfdtd_ref.so")
    src_time = max(os.path.getmtime(os.path.join(HERE, f))
                   for f in ("fdtd_oracle.c", "fdtd_oracle.h", "ref_shim.c", "Makefile"))
    stale = not os.path.exists(so) or os.path.getmtime(so) < src_time
    ref_missing = os.path.exists("/root/reference/main.c") and \
        (not os.path.exists(ref) or os.path.getmtime(ref) < src_time)
    if force or stale or ref_missing:
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


def restatement() -> Checker:
    if "o" not in _cache:
        build()
        _cache["o"] = Checker(C.CDLL(os.path.join(HERE, "libfdtd_oracle.so")), "oracle_", "port")
    return _cache["o"]


def reference():
    """The compiled reference, or None when oracle/_ref was never built."""
    if "r" not in _cache:
        path = os.path.join(HERE, "_ref", "libfdtd_ref.so")
        _cache["r"] = Checker(C.CDLL(path), "ref_", "reference") if os.path.exists(path) else None
    return _cache["r"]


def make_params(length, width, height, dx, dt, simulation_time, sampling_rate, mode) -> Params:
    """Derive the grid exactly like main.c:237-239 (float sizes, double step)."""
    p = Params()
    restatement()._make_params(length, width, height, dx, dt, simulation_time, sampling_rate, mode,
                               C.byref(p))
    return p


def step_count(p) -> int:
    return int(restatement()._step_count(C.byref(p)))


def source_bounds(p):
    b = (C.c_long * 4)()
    restatement()._source_bounds(C.byref(p), b)
    return tuple(int(x) for x in b)


def source_zte(p) -> float:
    return float(restatement()._source_zte(C.byref(p)))


def write_params(path, text_numbers):
    """Write a params.txt from the 8 number strings (no trailing newline, like the stock file)."""
    with open(path, "w") as fh:
        fh.write("\n".join(text_numbers))
    return path


STOCK_PARAMS = ("0.05", "0.05", "0.05", "0.001", "0.0000000000006", "0.00000000012", "2", "0")
