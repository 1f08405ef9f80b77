"""The reference's own geometry test, ported 1:1 (tests/test_geometry.jl:55-141): areas and volumes of every shape
class on every grid class, counted as (number of generated particles) x dr^dim, within the reference's tolerances
(1 % in 2-D, 3 % in 3-D).  Runs against the host generator (smoothedparticles.jl_b200/geometry.py), which restates
src/geometry.jl and src/grids.jl; the device generator is compared with the host generator point for point in
tests/test_generate_gpu.py.  Plus order checks of the lattice loops that the area/volume counts cannot see."""
import math

import numpy as np
import pytest

from smoothedparticles_jl_b200 import geometry as geo

RTOL_2D = 0.01   # test_geometry.jl:7-8
RTOL_3D = 0.03
N = 200          # :9
DA = 1 / (N * N)
DV = 1 / (N * N * N)


def rotmat(x):  # :47-53 (RealMatrix is column-major: the arguments are the columns)
    return np.array([[math.cos(x), -math.sin(x), 0.0], [math.sin(x), math.cos(x), 0.0], [0.0, 0.0, 1.0]])


def area(grid, shape):
    return len(geo.covering(grid, shape)) * DA


def volume(grid, shape):
    return len(geo.covering(grid, shape)) * DV


def test_area_tests():  # :56-104
    grid1, grid2, grid3 = geo.make_grid(1 / N, "square"), geo.make_grid(1 / N, "hexagonal"), geo.make_grid(1 / N, "vogel")
    s1 = geo.Circle(0.0, 0.0, 1.0)
    assert abs(area(grid1, s1) / math.pi - 1.0) < RTOL_2D
    s2 = geo.Rectangle(0.0, -1.0, 2.0, 5.0)
    assert abs(area(grid2, s2) / 12.0 - 1.0) < RTOL_2D
    s3 = geo.Ellipse(0.0, 0.0, 4.0, 1.0)
    assert abs(area(grid3, s3) / (4.0 * math.pi) - 1.0) < RTOL_2D
    tool1 = geo.Rectangle(0.0, -1.0, 4.0, 1.0)
    s4 = s3 - tool1
    assert abs(area(grid1, s4) / (2.0 * math.pi) - 1.0) < RTOL_2D
    s5 = s3 * tool1
    assert abs(area(grid2, s5) / (2.0 * math.pi) - 1.0) < RTOL_2D
    s6 = s4 + s5
    assert abs(area(grid3, s6) / (4.0 * math.pi) - 1.0) < RTOL_2D
    tool2 = geo.Rectangle(-4.0, -1.0, 4.0, 1.0)
    s7 = geo.Specification(tool2, lambda X: X[:, 1] < np.cos(math.pi * X[:, 0]))
    assert abs(area(grid1, s7) / 8.0 - 1.0) < RTOL_2D
    s8 = geo.Transform(s2, A=rotmat(math.pi / 7), b=(-2.0, 0.0, 0.0))
    assert abs(area(grid2, s8) / 12.0 - 1.0) < RTOL_2D
    s9 = geo.Polygon((-1.0, 0.0), (2.0, 0.0), (0.0, 3.0))
    assert abs(area(grid3, s9) / 4.5 - 1.0) < RTOL_2D


def test_volume_tests():  # :106-141
    grid1, grid2 = geo.make_grid(1 / N, "cubic"), geo.make_grid(1 / N, "facecentered")
    grid3, grid4 = geo.make_grid(1 / N, "bodycentered"), geo.make_grid(1 / N, "diamond")
    s1 = geo.Box(-0.7, -0.6, -0.5, 0.7, 0.6, 0.5)
    assert abs(volume(grid1, s1) / (1.4 * 1.2 * 1.0) - 1.0) < RTOL_3D
    s2 = geo.Ball(0.0, 0.0, 0.0, 0.8)
    assert abs(volume(grid2, s2) / (4 / 3 * math.pi * 0.8 ** 3) - 1.0) < RTOL_3D
    s3 = geo.Ellipsoid(0.0, 0.0, 0.0, 0.8, 0.5, 0.3)
    assert abs(volume(grid3, s3) / (4 / 3 * math.pi * 0.8 * 0.5 * 0.3) - 1.0) < RTOL_3D
    s4 = geo.Cone(0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.6, 0.3)
    assert abs(volume(grid4, s4) / (math.pi * (2 * 0.6 * 0.6 - 0.3 * 0.3) / 3) - 1.0) < RTOL_3D
    tool1 = geo.Polygon((0.0, 0.0), (0.6, 0.0), (0.0, 0.7))
    s5 = geo.RevolutionBody(tool1)
    assert abs(volume(grid1, s5) / (math.pi / 3 * 0.6 * 0.6 * 0.7) - 1.0) < RTOL_3D


# ---- order of generation (src/grids.jl: `for i in a, j in b, k in c` nests with the last range innermost; the
# centred points of the face-centred grid are pushed three per cell) against literal Python loops on small cases
def _ifloor(x):
    return int(math.floor(x))


def _iceil(x):
    return int(math.ceil(x))


def _ranges(box, a):
    return (range(_ifloor(box.x1_min / a), _iceil(box.x1_max / a) + 1), range(_ifloor(box.x2_min / a), _iceil(box.x2_max / a) + 1),
            range(_ifloor(box.x3_min / a), _iceil(box.x3_max / a) + 1))


def _inside(shape, x):
    return bool(shape.is_inside(np.array([x], dtype=np.float64))[0])


def test_lattice_orders_match_literal_loops():
    shape = geo.Ellipsoid(0.1, -0.05, 0.02, 0.31, 0.27, 0.22)
    box = shape.boundarybox()
    dr = 0.05
    # body-centred, grids.jl:150-175
    a = 2 ** (1 / 3) * dr
    ri, rj, rk = _ranges(box, a)
    want = [(i * a, j * a, k * a) for i in ri for j in rj for k in rk if _inside(shape, (i * a, j * a, k * a))]
    want += [((i + 0.5) * a, (j + 0.5) * a, (k + 0.5) * a) for i in ri for j in rj for k in rk
             if _inside(shape, ((i + 0.5) * a, (j + 0.5) * a, (k + 0.5) * a))]
    assert np.array_equal(geo.covering(geo.BodycenteredGrid(dr), shape), np.array(want))
    # face-centred, grids.jl:181-212
    a = 4 ** (1 / 3) * dr
    ri, rj, rk = _ranges(box, a)
    want = [(i * a, j * a, k * a) for i in ri for j in rj for k in rk if _inside(shape, (i * a, j * a, k * a))]
    for i in ri:
        for j in rj:
            for k in rk:
                for x in (((i + 0.5) * a, (j + 0.5) * a, k * a), ((i + 0.5) * a, j * a, (k + 0.5) * a),
                          (i * a, (j + 0.5) * a, (k + 0.5) * a)):
                    if _inside(shape, x):
                        want.append(x)
    assert np.array_equal(geo.covering(geo.FacecenteredGrid(dr), shape), np.array(want))
    # diamond, grids.jl:218-239
    a = 0.5 * dr
    ri, rj, rk = _ranges(box, a)
    want = []
    for i in ri:
        for j in rj:
            for k in rk:
                if (i % 2 != 0) == (j % 2 != 0) == (k % 2 != 0):
                    sm = int(math.fmod(i + j + k, 4))      # Julia's % is the remainder with the sign of the dividend
                    sm = (sm + 4) % 4
                    if sm in (0, 1) and _inside(shape, (i * a, j * a, k * a)):
                        want.append((i * a, j * a, k * a))
    assert np.array_equal(geo.covering(geo.DiamondGrid(dr), shape), np.array(want))
    # Vogel spiral, grids.jl:104-122
    disc = geo.Ellipse(0.2, 0.1, 0.5, 0.3)
    g = geo.VogelGrid(dr)
    bb = disc.boundarybox()
    R = max(math.sqrt(x * x + y * y) for x in (bb.x1_min, bb.x1_max) for y in (bb.x2_min, bb.x2_max))
    Nn = (R / g.k) ** 2
    want, n = [], 1.0
    while n <= Nn:
        x = (g.k * math.sqrt(n) * math.cos(n * geo.GOLDEN_ANGLE), g.k * math.sqrt(n) * math.sin(n * geo.GOLDEN_ANGLE), 0.0)
        if _inside(disc, x):
            want.append(x)
        n += 1.0
    got = geo.covering(g, disc)
    assert got.shape == (len(want), 3) and np.allclose(got, np.array(want), rtol=0, atol=1e-15)


def test_polygon_winding_and_closed_spline():
    tri = geo.Polygon((-1.0, 0.0), (2.0, 0.0), (0.0, 3.0))
    pts = np.array([[0.0, 1.0, 0.0], [1.9, 0.05, 0.0], [-1.0, 3.0, 0.0], [0.0, -0.1, 0.0], [0.0, 0.0, 0.0], [0.0, 3.0, 0.0]])
    # half-open in y (ys[i] <= y < ys[next]): the bottom edge belongs to the triangle, the apex does not
    assert tri.is_inside(pts).tolist() == [True, True, False, False, True, False]
    blob = geo.ClosedSpline((0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), n=64)
    assert blob.deg == 64 and abs(blob.xs[0] - blob.xs[-1]) < 1e-15 and abs(blob.ys[0] - blob.ys[-1]) < 1e-15
    inside = blob.is_inside(np.array([[0.5, 0.5, 0.0], [3.0, 3.0, 0.0]]))
    assert inside.tolist() == [True, False]


def test_device_generator_rejects_host_only_shapes():
    with pytest.raises(TypeError):
        geo.compile_shape(geo.Ellipse(0.0, 0.0, 1.0, 2.0))
    with pytest.raises(TypeError):
        geo.lattice_index_box(geo.DiamondGrid(0.1), geo.Ball(0.0, 0.0, 0.0, 1.0))
