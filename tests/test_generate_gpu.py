"""generate_particles! on the device (sp_generate.cu) against the host generator (geometry.py, the numpy restatement
of src/grids.jl + src/geometry.jl that the configs are built with): same particles, same Float64 coordinates, same
ORDER, for every grid / shape combination the BASELINE configs and the widened examples use."""
import numpy as np
import pytest

from smoothedparticles_jl_b200 import ParticleSystem, geometry as geo

pytestmark = pytest.mark.gpu


def _cases():
    dr = 5e-3
    cub = geo.CubicGrid(dr)
    box3 = geo.Box(0.0, 0.0, 0.0, 0.584, 0.35, 0.15)
    walls3 = geo.Specification(geo.BoundaryLayer(box3, cub, 2.5 * dr), geo.HalfSpace(1, "<", 0.35))   # collapse3d.jl:70-75
    yield "collapse3d fluid", cub, geo.Box(0.0, 0.0, 0.0, 0.142, 0.293, 0.15)
    yield "collapse3d walls", cub, walls3
    yield "ball", cub, geo.Ball(0.1, 0.05, 0.02, 0.0801)
    yield "ball minus box plus ball", cub, (geo.Ball(0.0, 0.0, 0.0, 0.06) - geo.Box(0.0, 0.0, 0.0, 1.0, 1.0, 1.0)) \
        + geo.Ball(0.05, 0.0, 0.0, 0.03)
    sq = geo.Squaregrid(1.5e-3)
    boxs = geo.Rectangle(0.0, 0.0, 0.14, 0.18)
    yield "static_container fluid", sq, geo.Rectangle(0.0, 0.0, 0.14, 0.14)
    yield "static_container walls", sq, geo.BoundaryLayer(boxs, sq, 2.5 * 1.5e-3)
    yield "collision disc", geo.Squaregrid(2e-2), geo.Circle(-0.5, -0.1, 0.4)
    hexg = geo.Hexagrid(0.01)
    cav = geo.Rectangle(0.0, 0.0, 1.0, 1.0)
    wall = geo.BoundaryLayer(cav, hexg, 0.03)
    yield "cavity fluid (hexagonal)", hexg, cav
    yield "cavity lid", hexg, geo.Specification(wall, geo.HalfSpace(1, ">", 1.0))       # cavity_flow.jl:62
    yield "cavity walls", hexg, geo.Specification(wall, geo.HalfSpace(1, "<=", 1.0))    # :63
    yield "intersection", hexg, geo.Circle(0.2, 0.2, 0.3) * geo.Rectangle(0.0, 0.0, 0.4, 0.25)
    yield "unit disc (test_IO.jl)", hexg, geo.Circle(0.0, 0.0, 1.0)


@pytest.mark.parametrize("name,grid,shape", list(_cases()), ids=[c[0] for c in _cases()])
def test_device_generation_is_the_host_generation(name, grid, shape):
    want = geo.covering(grid, shape)
    dom = geo.Box(-3.0, -3.0, -3.0, 3.0, 3.0, 3.0) if grid.dim == 3 else geo.Rectangle(-3.0, -3.0, 3.0, 3.0)
    s = ParticleSystem({"type": 1, "v": 3}, dom, 0.05)
    n = s.generate_particles(grid, shape, type=2.0)
    assert n == len(want) == len(s)
    assert np.array_equal(s.get("x"), want)                  # same doubles, same order
    assert np.all(s.get("type") == 2.0) and np.all(s.get("v") == 0.0)


def test_generation_appends_behind_existing_particles_in_any_device_order():
    sq = geo.Squaregrid(0.02)
    a, b = geo.Rectangle(0.0, 0.0, 0.5, 0.4), geo.Circle(1.0, 1.0, 0.3)
    s = ParticleSystem({"type": 1}, geo.Rectangle(-1.0, -1.0, 2.0, 2.0), 0.06)
    s.generate_particles(sq, a, type=0.0)
    rng = np.random.default_rng(0)
    s.set("x", s.get("x")[rng.permutation(len(s))])          # scramble, then sort by cell on the device
    xa = s.get("x")
    s.create_cell_list()
    s.generate_particles(sq, b, type=1.0)
    xb = geo.covering(sq, b)
    assert len(s) == len(xa) + len(xb)
    assert np.array_equal(s.get("x"), np.concatenate([xa, xb]))
    assert np.array_equal(s.get("type"), np.concatenate([np.zeros(len(xa)), np.ones(len(xb))]))
    s.create_cell_list()                                      # and the system is usable
    assert len(s) == len(xa) + len(xb)


def test_lambda_predicates_are_rejected_loudly():
    s = ParticleSystem({"type": 1}, geo.Rectangle(-1.0, -1.0, 2.0, 2.0), 0.06)
    with pytest.raises(TypeError):
        s.generate_particles(geo.Squaregrid(0.1), geo.Specification(geo.Rectangle(0, 0, 1, 1), lambda X: X[:, 0] < 0.5))


def test_collapse3d_initial_state_built_on_the_device():
    from smoothedparticles_jl_b200 import configs
    case = configs.collapse3d()
    a = case.make(ParticleSystem)
    b = case.make_on_device(ParticleSystem)
    assert len(a) == len(b) == case.n
    for nm in ("x", "v", "rho", "type", "P", "Dv", "Drho"):
        assert np.array_equal(a.get(nm), b.get(nm)), nm
    for _ in range(3):
        case.step(a)
        case.step(b)
    for nm in ("x", "v", "rho", "P"):
        assert np.array_equal(a.get(nm), b.get(nm)), nm
