"""tests/test_IO.jl of the reference, on the device path: a hexagonal covering of the unit circle with a scalar, a
vector and a matrix field is saved with save_frame, imported twice into a fresh system, and must come back exactly."""
import os

import numpy as np
import pytest

import smoothedparticles_jl_b200 as sp
from smoothedparticles_jl_b200 import ParticleSystem, geometry as geo, io as spio

pytestmark = pytest.mark.gpu


def _get_vars(x):  # tests/test_IO.jl:19-24
    s = x[:, 1].copy()
    v = np.stack([x[:, 1], -x[:, 0], np.zeros(len(x))], 1)
    M = x[:, :1] * np.arange(9.0)[None, :]
    return s, v, M


def _make_sys():  # :27-31
    return ParticleSystem({"s": 1, "v": 3, "M": 9}, geo.Circle(0.0, 0.0, 1.0), 0.1)


def test_save_and_import_round_trip(tmp_path):
    dr = 1 / 100
    sys_ = _make_sys()
    x = geo.covering(geo.Hexagrid(dr), geo.Circle(0.0, 0.0, 1.0))
    s, v, M = _get_vars(x)
    sys_.add_particles(x=x, s=s, v=v, M=M)
    sys_.create_cell_list()                      # the device order is now the cell order: output must not care
    out = spio.new_pvd_file(str(tmp_path / "test_IO"))
    spio.save_frame(out, sys_, "s", "v", "M")
    spio.save_pvd_file(out)
    assert os.path.exists(str(tmp_path / "test_IO" / "frame0.vtp"))       # "save data to vtk"
    assert os.path.exists(str(tmp_path / "test_IO" / "result.pvd"))
    new = _make_sys()
    spio.import_particles(new, str(tmp_path / "test_IO" / "frame0.vtp"))   # "read data from vtk"
    assert len(new) == len(sys_)
    for _ in range(2):
        xs = new.get("x")
        s2, v2, M2 = _get_vars(xs)
        assert np.array_equal(new.get("s"), s2) and np.array_equal(new.get("v"), v2) and np.array_equal(new.get("M"), M2)
        if len(new) == len(sys_):
            spio.import_particles(new, str(tmp_path / "test_IO" / "frame0.vtp"))   # "import even more particles"
            assert len(new) == 2 * len(sys_)


def test_import_keeps_the_constructor_defaults(tmp_path):
    # import_particles!(sys, path, constructor): fields the file does not carry keep what the constructor sets
    # (src/IO.jl; examples/cylinder.jl relies on it for rho0 and m)
    dr = 1 / 40
    src = _make_sys()
    x = geo.covering(geo.Hexagrid(dr), geo.Circle(0.0, 0.0, 1.0))
    s, v, M = _get_vars(x)
    src.add_particles(x=x, s=s, v=v, M=M)
    out = spio.new_pvd_file(str(tmp_path / "defaults"))
    spio.save_frame(out, src, "s", "v")                          # M is NOT in the file
    dst = ParticleSystem({"s": 1, "v": 3, "M": 9, "m": 1, "w": 3}, geo.Circle(0.0, 0.0, 1.0), 0.1)
    n = spio.import_particles(dst, str(tmp_path / "defaults" / "frame0.vtp"), m=2.5, w=(1.0, 2.0, 3.0), s=-7.0)
    assert n == len(x) == len(dst)
    assert np.array_equal(dst.get("s"), s) and np.array_equal(dst.get("v"), v)      # the file wins over the default
    assert np.all(dst.get("m") == 2.5) and np.all(dst.get("w") == np.array([1.0, 2.0, 3.0]))
    assert np.all(dst.get("M") == 0.0)
    n2 = spio.import_particles(dst, str(tmp_path / "defaults" / "frame0.vtp"),
                               constructor=lambda X: {"m": X[:, 0] ** 2, "M": np.eye(3).ravel()})
    assert np.array_equal(dst.get("m")[n:], dst.get("x")[n:, 0] ** 2)
    assert np.all(dst.get("M")[n:] == np.eye(3).ravel()) and len(dst) == n + n2
