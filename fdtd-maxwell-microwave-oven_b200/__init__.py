"""Thin ctypes view of libfdtd_b200.so (include/fdtd_b200.h) for the test-suite and bench.py.

The product is the C ABI plus the C host program; this module only marshals numpy arrays into
it.  There is NO fallback: if the shared library has not been built, importing fails, and every
device call raises FdtdError when CUDA is unavailable.  (The directory name carries hyphens, so
it is loaded through the `fdtd_b200` shim at the repository root.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDTD_B200_LIB") or os.path.join(HERE, "libfdtd_b200.so")  # override: A/B of two builds
PEER_BLOB_BYTES = 384  # FDTD_PEER_BLOB_BYTES
FIELD_NAMES = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
DUMP_NAMES = ("ex", "ey", "ez", "hx", "hy", "hz")

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {HERE}` "
                      "(or __graft_entry__.build()); there is no CPU fallback")

lib = C.CDLL(LIB_PATH)


class FdtdError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"fdtd_b200 error {code}: {message}")
        self.code = code


class Params(C.Structure):
    """fdtd_params == the reference's Parameters (main.c:57-71)."""
    _fields_ = [("length", C.c_float), ("width", C.c_float), ("height", C.c_float),
                ("spatial_step", C.c_double), ("time_step", C.c_double),
                ("simulation_time", C.c_float), ("sampling_rate", C.c_uint), ("mode", C.c_int),
                ("maxi", C.c_size_t), ("maxj", C.c_size_t), ("maxk", C.c_size_t)]

    def dims(self):
        return int(self.maxi), int(self.maxj), int(self.maxk)


class FieldPtrs(C.Structure):
    _fields_ = [(n, C.POINTER(C.c_double)) for n in FIELD_NAMES]


class SourcePlan(C.Structure):
    _fields_ = [("i0", C.c_long), ("i1", C.c_long), ("j0", C.c_long), ("j1", C.c_long),
                ("z_te", C.c_double)]


_BEGIN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.c_size_t)
_VARIABLE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_size_t)
_END = C.CFUNCTYPE(C.c_int, C.c_void_p)


class DumpSink(C.Structure):
    _fields_ = [("user", C.c_void_p), ("begin", _BEGIN), ("variable", _VARIABLE), ("end", _END)]


_D = C.POINTER(C.c_double)
_P = C.POINTER(Params)
_CTX = C.c_void_p

# every symbol include/fdtd_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "fdtd_abi_version": (C.c_int, []),
    "fdtd_last_error": (C.c_char_p, []),
    "fdtd_load_parameters": (C.c_int, [C.c_char_p, _P]),
    "fdtd_make_params": (C.c_int, [C.c_float, C.c_float, C.c_float, C.c_double, C.c_double,
                                   C.c_float, C.c_uint, C.c_int, _P]),
    "fdtd_field_sizes": (C.c_int, [_P, C.POINTER(C.c_size_t)]),
    "fdtd_step_count": (C.c_int, [_P, C.POINTER(C.c_size_t)]),
    "fdtd_source_plan_make": (C.c_int, [_P, C.POINTER(SourcePlan)]),
    "fdtd_source_values": (C.c_int, [_P, C.POINTER(SourcePlan), C.c_double, _D, _D]),
    "fdtd_initial_conditions_host": (C.c_int, [_P, _D]),
    "fdtd_slab_range": (C.c_int, [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_size_t)]),
    "fdtd_ctx_create": (C.c_int, [_P, C.c_int, C.POINTER(_CTX)]),
    "fdtd_ctx_create_slab": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(_CTX)]),
    "fdtd_ctx_destroy": (C.c_int, [_CTX]),
    "fdtd_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "fdtd_ctx_comm_init": (C.c_int, [_CTX, C.c_void_p]),
    "fdtd_ctx_peer_export": (C.c_int, [_CTX, C.c_void_p]),
    "fdtd_ctx_peer_connect": (C.c_int, [_CTX, C.c_void_p]),
    "fdtd_ctx_set_option": (C.c_int, [_CTX, C.c_char_p, C.c_long]),
    "fdtd_ctx_get_option": (C.c_int, [_CTX, C.c_char_p, C.POINTER(C.c_long)]),
    "fdtd_upload": (C.c_int, [_CTX, C.POINTER(FieldPtrs)]),
    "fdtd_download": (C.c_int, [_CTX, C.POINTER(FieldPtrs)]),
    "fdtd_upload_slab": (C.c_int, [_CTX, C.POINTER(FieldPtrs)]),
    "fdtd_download_slab": (C.c_int, [_CTX, C.POINTER(FieldPtrs)]),
    "fdtd_set_initial_conditions": (C.c_int, [_CTX]),
    "fdtd_set_source": (C.c_int, [_CTX, C.c_double]),
    "fdtd_update_H_field": (C.c_int, [_CTX]),
    "fdtd_update_E_field": (C.c_int, [_CTX]),
    "fdtd_run": (C.c_int, [_CTX, C.c_size_t, _D]),
    "fdtd_run_timed": (C.c_int, [_CTX, C.c_size_t, _D, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float)]),
    "fdtd_sync": (C.c_int, [_CTX]),
    "fdtd_run_hosted": (C.c_int, [_CTX, C.POINTER(FieldPtrs), C.c_size_t, _D]),
    "fdtd_aggregate": (C.c_int, [_CTX, C.c_int, _D]),
    "fdtd_propagate": (C.c_int, [_CTX, C.POINTER(DumpSink), C.POINTER(C.c_size_t), _D]),
    "fdtd_group_create": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "fdtd_group_create_transport": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "fdtd_group_destroy": (C.c_int, [C.c_void_p]),
    "fdtd_group_size": (C.c_int, [C.c_void_p]),
    "fdtd_group_ctx": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_CTX)]),
    "fdtd_group_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_long]),
    "fdtd_group_upload": (C.c_int, [C.c_void_p, C.POINTER(FieldPtrs)]),
    "fdtd_group_download": (C.c_int, [C.c_void_p, C.POINTER(FieldPtrs)]),
    "fdtd_group_set_initial_conditions": (C.c_int, [C.c_void_p]),
    "fdtd_group_run": (C.c_int, [C.c_void_p, C.c_size_t, _D]),
    "fdtd_group_sync": (C.c_int, [C.c_void_p]),
    "fdtd_group_aggregate": (C.c_int, [C.c_void_p, C.c_int, _D]),
    "fdtd_group_energy": (C.c_int, [C.c_void_p, C.c_int, _D, _D]),
    "fdtd_group_propagate": (C.c_int, [C.c_void_p, C.POINTER(DumpSink), C.POINTER(C.c_size_t), _D]),
    "fdtd_energy": (C.c_int, [_CTX, C.c_int, _D, _D]),
    "fdtd_validation_error": (C.c_int, [_CTX, C.c_double, _D, _D]),
    "fdtd_fill_test_pattern": (C.c_int, [_CTX, C.c_ulonglong]),
    "fdtd_checksum": (C.c_int, [_CTX, C.POINTER(C.c_ulonglong)]),
    "fdtd_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "fdtd_host_alloc_near": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "fdtd_host_numa_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fdtd_host_free": (C.c_int, [C.c_void_p]),
    "fdtd_ctx_info": (C.c_int, [_CTX, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype, _fn.argtypes = _res, _args


def _check(rc):
    if rc != 0:
        raise FdtdError(rc, (lib.fdtd_last_error() or b"").decode(errors="replace"))


# ---- host-side helpers ---------------------------------------------------------------------

def load_parameters(path) -> Params:
    p = Params()
    _check(lib.fdtd_load_parameters(os.fsencode(path), C.byref(p)))
    return p


def make_params(length, width, height, spatial_step, time_step, simulation_time, sampling_rate,
                mode) -> Params:
    p = Params()
    _check(lib.fdtd_make_params(length, width, height, spatial_step, time_step, simulation_time,
                                sampling_rate, mode, C.byref(p)))
    return p


def field_sizes(p):
    out = (C.c_size_t * 6)()
    _check(lib.fdtd_field_sizes(C.byref(p), out))
    return [int(x) for x in out]


def field_shapes(p):
    """(planes, rows, row length) of the six dense host arrays (main.c:379-407)."""
    nx, ny, nz = p.dims()
    return {"Ex": (nz + 1, ny + 1, nx), "Ey": (nz + 1, ny, nx + 1), "Ez": (nz, ny + 1, nx + 1),
            "Hx": (nz, ny, nx + 1), "Hy": (nz, ny + 1, nx), "Hz": (nz + 1, ny, nx)}


def step_count(p) -> int:
    n = C.c_size_t()
    _check(lib.fdtd_step_count(C.byref(p), C.byref(n)))
    return int(n.value)


def source_plan(p) -> SourcePlan:
    plan = SourcePlan()
    _check(lib.fdtd_source_plan_make(C.byref(p), C.byref(plan)))
    return plan


def source_values(p, plan, t):
    n = int(plan.i1 - plan.i0)
    ez, hx = np.empty(n), np.empty(n)
    _check(lib.fdtd_source_values(C.byref(p), C.byref(plan), float(t), ez.ctypes.data_as(_D),
                                  hx.ctypes.data_as(_D)))
    return ez, hx


def initial_conditions_host(p):
    ey = np.empty(field_shapes(p)["Ey"])
    _check(lib.fdtd_initial_conditions_host(C.byref(p), ey.ctypes.data_as(_D)))
    return ey


def slab_range(maxk, rank, nranks):
    k0, k1 = C.c_size_t(), C.c_size_t()
    _check(lib.fdtd_slab_range(maxk, rank, nranks, C.byref(k0), C.byref(k1)))
    return int(k0.value), int(k1.value)


def host_numa_info(device=0):
    nodes, node = C.c_int(), C.c_int()
    _check(lib.fdtd_host_numa_info(int(device), C.byref(nodes), C.byref(node)))
    return {"nodes": int(nodes.value), "node_of_device": int(node.value)}


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib.fdtd_nccl_unique_id(buf))
    return buf.raw


# The halo plan implemented by exchange_h / exchange_e in csrc/fdtd_ctx.cu, spelled
# out so the CPU tests can run the same plan over gloo with the oracle as the per-slab kernel:
# after each half-step one plane of two components travels to one neighbour.
HALO_PLAN = {
    # after update_H_field: my top owned CELL plane (k1-1) of Hx, Hy goes to rank+1, which keeps it
    # as its plane k0-1 (read by its Ex/Ey update at node plane k0, main.c:486-493)
    "after_H": {"fields": ("Hx", "Hy"), "send_plane": "k1-1", "to": +1, "recv_plane": "k0-1"},
    # after update_E_field: my first owned NODE plane (k0) of Ex, Ey goes to rank-1, which keeps it
    # as its plane k1 (read by its Hx/Hy update at cell plane k1-1, main.c:448-455)
    "after_E": {"fields": ("Ex", "Ey"), "send_plane": "k0", "to": -1, "recv_plane": "k1"},
}
# The fused step (kernels 2 and 3) exchanges once per time step: the slab above recomputes H_new of
# the plane below its first one, so it needs that plane's Ex, Ey, Ez as well as Hx, Hy.
HALO_PLAN_FUSED = {
    "after_step_up": {"fields": ("Hx", "Hy", "Ex", "Ey", "Ez"), "send_plane": "k1-1", "to": +1, "recv_plane": "k0-1"},
    "after_step_down": {"fields": ("Ex", "Ey"), "send_plane": "k0", "to": -1, "recv_plane": "k1"},
}


class PinnedArrays:
    """Six pinned host arrays in the reference's dense layout (fdtd_host_alloc)."""

    def __init__(self, p, fill=None, shapes=None, device=None):
        """device: allocate near that GPU (fdtd_host_alloc_near)"""
        self._ptrs = []
        self.arrays = {}
        for name, shape in (shapes or field_shapes(p)).items():
            n = int(np.prod(shape))
            ptr = C.c_void_p()
            if device is None:
                _check(lib.fdtd_host_alloc(max(n, 1) * 8, C.byref(ptr)))
            else:
                _check(lib.fdtd_host_alloc_near(int(device), max(n, 1) * 8, C.byref(ptr)))
            self._ptrs.append(ptr)
            buf = (C.c_double * n).from_address(ptr.value)
            arr = np.frombuffer(buf, dtype=np.float64).reshape(shape)
            if fill is not None:
                arr[...] = fill
            self.arrays[name] = arr

    def close(self):
        self.arrays = {}
        for ptr in self._ptrs:
            lib.fdtd_host_free(ptr)
        self._ptrs = []


def _ptrs(fields):
    fp = FieldPtrs()
    for n in FIELD_NAMES:
        a = fields[n]
        if a.dtype != np.float64 or not a.flags.c_contiguous:
            raise ValueError(f"{n}: need a C-contiguous float64 array")
        setattr(fp, n, a.ctypes.data_as(_D))
    return fp


class Context:
    """fdtd_ctx: the six arrays resident in HBM plus streams (include/fdtd_b200.h)."""

    def __init__(self, p: Params, device=0, rank=0, nranks=1):
        self.p = p
        self._h = _CTX()
        if nranks == 1:
            _check(lib.fdtd_ctx_create(C.byref(p), device, C.byref(self._h)))
        else:
            _check(lib.fdtd_ctx_create_slab(C.byref(p), device, rank, nranks, C.byref(self._h)))
        self.rank, self.nranks = rank, nranks
        self.k0, self.k1 = slab_range(int(p.maxk), rank, nranks)

    def close(self):
        if self._h:
            lib.fdtd_ctx_destroy(self._h)
            self._h = _CTX()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, unique_id: bytes):
        _check(lib.fdtd_ctx_comm_init(self._h, C.create_string_buffer(unique_id, 128)))

    def peer_export(self) -> bytes:
        """descriptor of this slab's state for fdtd_ctx_peer_connect on the other ranks"""
        buf = C.create_string_buffer(PEER_BLOB_BYTES)
        _check(lib.fdtd_ctx_peer_export(self._h, buf))
        return buf.raw

    def peer_connect(self, blobs):
        """blobs: the peer_export() of every rank, in rank order"""
        table = b"".join(blobs)
        assert len(table) == PEER_BLOB_BYTES * self.nranks
        _check(lib.fdtd_ctx_peer_connect(self._h, C.create_string_buffer(table, len(table))))

    def set_option(self, key, value):
        _check(lib.fdtd_ctx_set_option(self._h, key.encode(), int(value)))

    def get_option(self, key):
        v = C.c_long()
        _check(lib.fdtd_ctx_get_option(self._h, key.encode(), C.byref(v)))
        return int(v.value)

    def info(self):
        a, b, c_, d = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        _check(lib.fdtd_ctx_info(self._h, C.byref(a), C.byref(b), C.byref(c_), C.byref(d)))
        return {"hbm_bytes": int(a.value), "pitch": int(b.value), "rows": int(c_.value),
                "planes": int(d.value)}

    def upload(self, fields):
        _check(lib.fdtd_upload(self._h, C.byref(_ptrs(fields))))

    def download(self, out=None):
        if out is None:
            out = {n: np.zeros(s) for n, s in field_shapes(self.p).items()}
        _check(lib.fdtd_download(self._h, C.byref(_ptrs(out))))
        return out

    def slab_shapes(self):
        """dense shapes of the arrays holding only this slab's planes (fdtd_upload_slab)."""
        nk = self.k1 - self.k0
        top = 1 if self.rank == self.nranks - 1 else 0
        out = {}
        for name, (planes, rows, cols) in field_shapes(self.p).items():
            out[name] = ((nk + top) if name in ("Ex", "Ey", "Hz") else nk, rows, cols)
        return out

    def upload_slab(self, fields):
        _check(lib.fdtd_upload_slab(self._h, C.byref(_ptrs(fields))))

    def download_slab(self, out):
        _check(lib.fdtd_download_slab(self._h, C.byref(_ptrs(out))))
        return out

    def set_initial_conditions(self):
        _check(lib.fdtd_set_initial_conditions(self._h))

    def set_source(self, t):
        _check(lib.fdtd_set_source(self._h, float(t)))

    def update_H_field(self):
        _check(lib.fdtd_update_H_field(self._h))

    def update_E_field(self):
        _check(lib.fdtd_update_E_field(self._h))

    def run(self, steps, t=0.0):
        tc = C.c_double(t)
        _check(lib.fdtd_run(self._h, int(steps), C.byref(tc)))
        return tc.value

    def run_timed(self, steps, t=0.0, per_kernel=True):
        tc = C.c_double(t)
        total, h, e = C.c_float(), C.c_float(), C.c_float()
        _check(lib.fdtd_run_timed(self._h, int(steps), C.byref(tc), C.byref(total),
                                  C.byref(h) if per_kernel else None,
                                  C.byref(e) if per_kernel else None))
        return tc.value, total.value, h.value, e.value

    def sync(self):
        _check(lib.fdtd_sync(self._h))

    def run_hosted(self, fields, steps, t=0.0):
        """advance host arrays (this slab's planes, dense layout) in place by `steps` time steps"""
        tc = C.c_double(t)
        _check(lib.fdtd_run_hosted(self._h, C.byref(_ptrs(fields)), int(steps), C.byref(tc)))
        return tc.value

    def aggregate(self, var):
        nx, ny, _ = self.p.dims()
        out = np.empty((self.k1 - self.k0, ny, nx))
        _check(lib.fdtd_aggregate(self._h, int(var), out.ctypes.data_as(_D)))
        return out

    def energy(self, as_coded=False):
        e, h = C.c_double(), C.c_double()
        _check(lib.fdtd_energy(self._h, 1 if as_coded else 0, C.byref(e), C.byref(h)))
        return e.value, h.value

    def validation_error(self, t):
        sums, rel = (C.c_double * 6)(), (C.c_double * 3)()
        _check(lib.fdtd_validation_error(self._h, float(t), sums, rel))
        return list(sums), list(rel)

    def fill_test_pattern(self, seed):
        _check(lib.fdtd_fill_test_pattern(self._h, int(seed)))

    def checksum(self):
        out = (C.c_ulonglong * 6)()
        _check(lib.fdtd_checksum(self._h, out))
        return [int(x) for x in out]

    def propagate(self, on_begin=None, on_variable=None, on_end=None, dumps=True):
        """propagate_fields (main.c:755-799).  Callbacks run on the library's writer thread."""
        steps, tc = C.c_size_t(), C.c_double()
        if not dumps:
            _check(lib.fdtd_propagate(self._h, None, C.byref(steps), C.byref(tc)))
            return int(steps.value), tc.value

        def _begin(user, iteration, dims, k0):
            if on_begin:
                on_begin(iteration, (dims[0], dims[1], dims[2]), k0)
            return 0

        def _variable(user, name, data, count):
            if on_variable:
                on_variable(name.decode(), np.ctypeslib.as_array(data, shape=(count,)).copy())
            return 0

        def _end(user):
            if on_end:
                on_end()
            return 0

        sink = DumpSink(None, _BEGIN(_begin), _VARIABLE(_variable), _END(_end))
        _check(lib.fdtd_propagate(self._h, C.byref(sink), C.byref(steps), C.byref(tc)))
        return int(steps.value), tc.value


class Group:
    """fdtd_group: every z-slab of the cavity driven from this one thread (no launcher)."""

    def __init__(self, p: Params, ngpus, devices=None, transport=None):
        """transport: None (FDTD_B200_TRANSPORT or peer copies), "peer" or "nccl"; devices may repeat
        (several slabs on one GPU) with peer copies"""
        self.p = p
        self._g = C.c_void_p()
        dev = (C.c_int * ngpus)(*devices) if devices is not None else None
        if transport is None:
            _check(lib.fdtd_group_create(C.byref(p), ngpus, dev, C.byref(self._g)))
        else:
            _check(lib.fdtd_group_create_transport(C.byref(p), ngpus, dev, {"peer": 0, "nccl": 1}[transport],
                                                   C.byref(self._g)))
        self.n = lib.fdtd_group_size(self._g)
        self.slabs = []
        for r in range(self.n):
            h = _CTX()
            _check(lib.fdtd_group_ctx(self._g, r, C.byref(h)))
            ctx = Context.__new__(Context)      # a view: the group owns the context
            ctx.p, ctx._h, ctx.rank, ctx.nranks = p, h, r, self.n
            ctx.k0, ctx.k1 = slab_range(int(p.maxk), r, self.n)
            ctx.close = lambda: None
            self.slabs.append(ctx)

    def close(self):
        if self._g:
            for s in self.slabs:
                s._h = _CTX()
            lib.fdtd_group_destroy(self._g)
            self._g = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_option(self, key, value):
        _check(lib.fdtd_group_set_option(self._g, key.encode(), int(value)))

    def upload(self, fields):
        _check(lib.fdtd_group_upload(self._g, C.byref(_ptrs(fields))))

    def download(self, out=None):
        if out is None:
            out = {n: np.zeros(s) for n, s in field_shapes(self.p).items()}
        _check(lib.fdtd_group_download(self._g, C.byref(_ptrs(out))))
        return out

    def set_initial_conditions(self):
        _check(lib.fdtd_group_set_initial_conditions(self._g))

    def run(self, steps, t=0.0):
        tc = C.c_double(t)
        _check(lib.fdtd_group_run(self._g, int(steps), C.byref(tc)))
        return tc.value

    def sync(self):
        _check(lib.fdtd_group_sync(self._g))

    def aggregate(self, var):
        nx, ny, nz = self.p.dims()
        out = np.empty((nz, ny, nx))
        _check(lib.fdtd_group_aggregate(self._g, int(var), out.ctypes.data_as(_D)))
        return out

    def energy(self, as_coded=False):
        e, h = C.c_double(), C.c_double()
        _check(lib.fdtd_group_energy(self._g, 1 if as_coded else 0, C.byref(e), C.byref(h)))
        return e.value, h.value

    def propagate(self, on_begin=None, on_variable=None, on_end=None, dumps=True):
        """callbacks get the slab rank as first argument; they run on per-slab writer threads"""
        steps, tc = C.c_size_t(), C.c_double()
        if not dumps:
            _check(lib.fdtd_group_propagate(self._g, None, C.byref(steps), C.byref(tc)))
            return int(steps.value), tc.value
        keep, sinks = [], (DumpSink * self.n)()
        for r in range(self.n):
            def _begin(user, iteration, dims, k0, r=r):
                if on_begin:
                    on_begin(r, iteration, (dims[0], dims[1], dims[2]), k0)
                return 0

            def _variable(user, name, data, count, r=r):
                if on_variable:
                    on_variable(r, name.decode(), np.ctypeslib.as_array(data, shape=(count,)).copy())
                return 0

            def _end(user, r=r):
                if on_end:
                    on_end(r)
                return 0
            cbs = (_BEGIN(_begin), _VARIABLE(_variable), _END(_end))
            keep.append(cbs)
            sinks[r] = DumpSink(None, *cbs)
        _check(lib.fdtd_group_propagate(self._g, sinks, C.byref(steps), C.byref(tc)))
        return int(steps.value), tc.value


# ---- numpy mirrors of the test pattern / checksum (tests compare the device against these) -----

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def pattern_host(p, seed):
    """What fdtd_fill_test_pattern writes, as six dense numpy arrays."""
    out = {}
    for a, (name, shape) in enumerate(field_shapes(p).items()):
        dense = np.arange(int(np.prod(shape)), dtype=np.uint64)
        r = _splitmix64(np.uint64(seed) ^ (np.uint64(a) << np.uint64(58)) ^ dense)
        u = (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        out[name] = (2.0 * u - 1.0).reshape(shape)
    return out


def checksum_host(fields):
    """What fdtd_checksum returns for a single-GPU context holding `fields`."""
    out = []
    for name in FIELD_NAMES:
        a = np.ascontiguousarray(fields[name])
        dense = np.arange(a.size, dtype=np.uint64)
        with np.errstate(over="ignore"):
            s = _splitmix64(a.reshape(-1).view(np.uint64) + dense)
            out.append(int(np.add.reduce(s, dtype=np.uint64)))
    return out
