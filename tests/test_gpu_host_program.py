"""The C host program (host/microwave.c) end to end on a GPU: same console output and exit codes as
the reference's main() (main.c:807-853), same dump cadence and dump contents as its
propagate_fields / write_silo (golden fixture generated from the compiled reference)."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from pdb_reader import SiloFile

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "fdtd-maxwell-microwave-oven_b200", "microwave")

STDOUT_HEAD = ["Welcome into our microwave oven eletrico-magnetic field simulator! ",
               "Loading the parameters...", "Initializing fields"]
STDOUT_TAIL = ["Creating mesh", "Setting initial conditions", "Launching simulation", "Freeing memory...",
               "Simulation complete!"]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def exe(F):
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build()
    return EXE


@pytest.mark.parametrize("mode", [0, 1])
def test_microwave_writes_the_reference_silo_files(exe, golden, tmp_path, mode):
    """./microwave params.txt writes r/result%04d.silo (PDB driver layout) with the mesh, the variables and
    the defvars of write_silo(), main.c:550-598; parsed back with the independent reader, every variable
    has the sha256 of the reference's own dump."""
    g = golden["propagate_tiny"][f"mode{mode}"]
    (tmp_path / "params.txt").write_text("\n".join(g["params"]))
    (tmp_path / "r").mkdir()
    r = subprocess.run([exe, "params.txt"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    want = STDOUT_HEAD + (["Validation mode activated. "] if mode == 0 else []) + STDOUT_TAIL
    assert r.stdout.splitlines() == want
    nx, ny, nz = g["grid"]
    names = ["ex", "ey", "ez", "hx", "hy", "hz"] + (["aEy", "aHx", "aHz"] if mode == 0 else [])
    files = sorted(p.name for p in (tmp_path / "r").iterdir())
    assert files == [os.path.basename(d["file"]) for d in g["dumps"]]       # result0001.silo, ... and nothing else
    dx = float(g["params"][3])
    for d in g["dumps"]:
        s = SiloFile(tmp_path / d["file"])
        assert s.objects() == ["mesh"] + names + ["vecs"]
        mesh = s.object("mesh")
        assert mesh["_type"] == "quadmesh" and list(mesh["dims"]) == [nx + 1, ny + 1, nz + 1]   # main.c:257-259
        for axis, n in enumerate((nx, ny, nz)):
            assert np.array_equal(mesh[f"coord{axis}"], np.arange(n + 1) * dx)                  # main.c:270-278
        for name in names:
            v = s.object(name)
            assert v["_type"] == "quadvar" and v["meshid"] == "mesh" and list(v["dims"]) == [nx, ny, nz]
            assert v["centering"] == 111 and v["datatype"] == 20                                # DB_ZONECENT, DB_DOUBLE
            assert digest(v["value0"]) == d["vars"][name], (d["file"], name)
        vecs = s.object("vecs")
        assert vecs["names"] == ";E;H" and vecs["defns"] == ";{ex, ey, ez};{hx, hy, hz}"          # main.c:591-593


@pytest.mark.parametrize("mode", [0, 1])
def test_microwave_raw_sink(exe, golden, tmp_path, mode):
    """FDTD_B200_SINK=raw: the same variables as one raw brick plus a VisIt BOV header per variable"""
    g = golden["propagate_tiny"][f"mode{mode}"]
    (tmp_path / "params.txt").write_text("\n".join(g["params"]))
    (tmp_path / "r").mkdir()
    r = subprocess.run([exe, "params.txt"], cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, FDTD_B200_SINK="raw"))
    assert r.returncode == 0, r.stderr
    nx, ny, nz = g["grid"]
    n = nx * ny * nz
    names = ["ex", "ey", "ez", "hx", "hy", "hz"] + (["aEy", "aHx", "aHz"] if mode == 0 else [])
    raws = sorted(p.name for p in (tmp_path / "r").glob("*.raw"))
    assert raws == [os.path.basename(d["file"]).replace(".silo", ".raw") for d in g["dumps"]]
    for d in g["dumps"]:
        base = os.path.basename(d["file"]).replace(".silo", "")
        data = np.fromfile(tmp_path / "r" / (base + ".raw"))
        assert data.size == n * len(names)
        for v, name in enumerate(names):
            assert digest(data[v * n:(v + 1) * n]) == d["vars"][name], (base, name)
            bov = (tmp_path / "r" / f"{base}.{name}.bov").read_text()
            assert f"DATA_SIZE: {nx} {ny} {nz}" in bov and f"BYTE_OFFSET: {v * n * 8}" in bov


def test_microwave_error_paths(exe, tmp_path):
    # wrong argument count: perror text of main.c:813, exit status EXIT_FAILURE
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "This program needs 1 argument" in r.stderr
    r = subprocess.run([exe, "nope.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "Unable to open parameters file!" in r.stderr
    # time step larger than the simulated time (main.c:818-821)
    (tmp_path / "p.txt").write_text("0.012\n0.011\n0.010\n0.001\n0.0000000000006\n0.0000000000001\n3\n1")
    r = subprocess.run([exe, "p.txt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "The time step must be lower than the simulation time!" in r.stderr
    # no r/ directory: the reference dies in write_silo with "Could not create DB" (main.c:556-559)
    (tmp_path / "q.txt").write_text("0.012\n0.011\n0.010\n0.001\n0.0000000000006\n0.000000000012\n3\n1")
    r = subprocess.run([exe, "q.txt"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Could not create DB" in r.stderr
