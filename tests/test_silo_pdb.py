"""The minimal Silo/PDB writer (host/silo_pdb.c, what `microwave` writes r/result%04d.silo with) against
an independent reader of the PDB container (tests/pdb_reader.py): the objects write_silo() creates
(main.c:550-598) come back with the right names, shapes, types and bit-identical data."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from pdb_reader import PDBError, SiloFile

LIB = os.path.join(ROOT, "fdtd-maxwell-microwave-oven_b200", "libsilo_pdb.so")


@pytest.fixture(scope="module")
def spdb():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    lib = C.CDLL(LIB)
    D, I = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.spdb_create.restype = C.c_void_p
    lib.spdb_create.argtypes = [C.c_char_p, C.c_char_p]
    lib.spdb_put_quadmesh.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(D), I]
    lib.spdb_quadvar_begin.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, I]
    lib.spdb_quadvar_append.argtypes = [C.c_void_p, D, C.c_size_t]
    lib.spdb_quadvar_end.argtypes = [C.c_void_p]
    lib.spdb_put_defvars.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), I, C.POINTER(C.c_char_p)]
    lib.spdb_put_multimesh.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p)]
    lib.spdb_put_multivar.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p)]
    lib.spdb_close.argtypes = [C.c_void_p]
    return lib


def _write(lib, path, dims, variables, pieces=1):
    nx, ny, nz = dims
    D = C.POINTER(C.c_double)
    f = lib.spdb_create(os.fsencode(path), None)
    assert f
    coords = [np.arange(n + 1) * 0.001 for n in dims]
    cp = (D * 3)(*[c.ctypes.data_as(D) for c in coords])
    assert lib.spdb_put_quadmesh(f, b"mesh", cp, (C.c_int * 3)(nx + 1, ny + 1, nz + 1)) == 0
    for name, data in variables.items():
        assert lib.spdb_quadvar_begin(f, name.encode(), b"mesh", (C.c_int * 3)(nx, ny, nz)) == 0
        flat = np.ascontiguousarray(data).reshape(-1)
        for part in np.array_split(flat, pieces):
            part = np.ascontiguousarray(part)
            assert lib.spdb_quadvar_append(f, part.ctypes.data_as(D), part.size) == 0
        assert lib.spdb_quadvar_end(f) == 0
    names = (C.c_char_p * 2)(b"E", b"H")
    defs = (C.c_char_p * 2)(b"{ex, ey, ez}", b"{hx, hy, hz}")
    assert lib.spdb_put_defvars(f, b"vecs", 2, names, (C.c_int * 2)(201, 201), defs) == 0
    assert lib.spdb_close(f) == 0
    return coords


@pytest.mark.parametrize("dims,pieces", [((5, 4, 3), 1), ((17, 9, 11), 3), ((1, 1, 1), 1)])
def test_write_silo_layout_round_trip(spdb, tmp_path, dims, pieces):
    rng = np.random.default_rng(3)
    nx, ny, nz = dims
    variables = {n: rng.uniform(-1, 1, size=(nz, ny, nx)) for n in ("ex", "ey", "ez", "hx", "hy", "hz", "aEy", "aHx", "aHz")}
    path = tmp_path / "result0001.silo"
    coords = _write(spdb, path, dims, variables, pieces)
    s = SiloFile(path)
    assert s.extras["Major-Order"] == b"101" and s.extras["Has-Directories"] == b"1"
    assert s.symbols["/"]["type"] == "Directory"
    assert s.objects() == ["mesh"] + list(variables) + ["vecs"]          # the order write_silo() creates them in
    mesh = s.object("mesh")
    assert mesh["_type"] == "quadmesh" and mesh["ndims"] == 3 and mesh["nspace"] == 3
    assert mesh["coordtype"] == 130 and mesh["datatype"] == 20 and mesh["major_order"] == 0   # DB_COLLINEAR, DB_DOUBLE
    assert list(mesh["dims"]) == [nx + 1, ny + 1, nz + 1] and mesh["nnodes"] == (nx + 1) * (ny + 1) * (nz + 1)
    for d in range(3):
        got = mesh[f"coord{d}"]
        assert got.dtype == np.float64 and np.array_equal(got.view(np.uint64), coords[d].view(np.uint64))
    assert list(mesh["min_extents"]) == [0.0, 0.0, 0.0]
    assert list(mesh["max_extents"]) == [coords[d][-1] for d in range(3)]
    for name, want in variables.items():
        v = s.object(name)
        assert v["_type"] == "quadvar" and v["meshid"] == "mesh" and v["centering"] == 111    # DB_ZONECENT
        assert v["datatype"] == 20 and v["nvals"] == 1 and v["nels"] == nx * ny * nz
        assert list(v["dims"]) == [nx, ny, nz] and list(v["align"]) == [0.5, 0.5, 0.5]
        got = v["value0"]
        assert got.dtype == np.float64 and got.size == want.size
        assert np.array_equal(got.view(np.uint64), want.reshape(-1).view(np.uint64))         # x fastest
    vecs = s.object("vecs")
    assert vecs["_type"] == "defvars" and vecs["ndefs"] == 2
    assert vecs["names"] == ";E;H" and vecs["defns"] == ";{ex, ey, ez};{hx, hy, hz}" and list(vecs["types"]) == [201, 201]


def test_multiblock_root_file(spdb, tmp_path):
    path = tmp_path / "root.silo"
    f = spdb.spdb_create(os.fsencode(path), b"two slabs")
    blocks = (C.c_char_p * 2)(b"result0001.slab0.silo:/mesh", b"result0001.slab1.silo:/mesh")
    assert spdb.spdb_put_multimesh(f, b"mesh", 2, blocks) == 0
    vblocks = (C.c_char_p * 2)(b"result0001.slab0.silo:/ex", b"result0001.slab1.silo:/ex")
    assert spdb.spdb_put_multivar(f, b"ex", 2, vblocks) == 0
    assert spdb.spdb_close(f) == 0
    s = SiloFile(path)
    assert s.read_string("/_fileinfo") == "two slabs"
    mm, mv = s.object("mesh"), s.object("ex")
    assert mm["_type"] == "multimesh" and mm["nblocks"] == 2 and list(mm["meshtypes"]) == [130, 130]
    assert mm["meshnames"] == ";result0001.slab0.silo:/mesh;result0001.slab1.silo:/mesh"
    assert mv["_type"] == "multivar" and mv["varnames"].split(";")[1:] == [b.decode() for b in vblocks]


def test_reader_rejects_damage(spdb, tmp_path):
    """the reader is a real parser: a truncated file or a wrong struct size is an error, not silence"""
    path = tmp_path / "x.silo"
    _write(spdb, path, (3, 3, 3), {"ex": np.zeros((3, 3, 3))})
    raw = open(path, "rb").read()
    bad = tmp_path / "bad.silo"
    bad.write_bytes(raw.replace(b"Group\x0140\x01", b"Group\x0148\x01"))
    with pytest.raises(PDBError):
        SiloFile(bad).object("mesh")
    bad.write_bytes(raw[:len(raw) // 2])
    with pytest.raises((PDBError, ValueError, IndexError)):
        SiloFile(bad)


def test_create_fails_without_directory(spdb, tmp_path):
    assert not spdb.spdb_create(os.fsencode(tmp_path / "missing" / "result0001.silo"), None)
