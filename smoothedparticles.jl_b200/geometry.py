"""Host-side shapes and lattice generators (set-up only, never on the step loop).

Vectorised numpy restatement of the few pieces of the reference's set-up code needed to
regenerate the initial states of the BASELINE configs with identical Float64 positions and
identical particle ORDER:

* shapes: ``Box``/``Rectangle``/``Circle``/``Ball``, boolean ``+ - *``, ``Specification``,
  ``BoundaryLayer``  — reference ``src/geometry.py`` counterpart: ``src/geometry.jl:15-43, 49-68,
  108-234, 240-258``
* grids: ``Squaregrid``, ``Hexagrid``, ``CubicGrid`` and ``covering`` — ``src/grids.jl:48-91, 124-144``
* ``generate_positions`` = the position part of ``generate_particles!`` — ``src/grids.jl:253-258``
* the rest of the reference's geometry module, host side only (the device generator takes the subset above and fails
  loudly otherwise): ``Ellipse``, ``Ellipsoid``, ``Transform``, ``Polygon``, ``ClosedSpline``, ``Cone``,
  ``RevolutionBody`` — ``src/geometry.jl:70-100, 262-440``; ``VogelGrid``, ``BodycenteredGrid``, ``FacecenteredGrid``,
  ``DiamondGrid`` — ``src/grids.jl:93-251``.  Pinned by the reference's own ``tests/test_geometry.jl``
  (``tests/test_geometry_pins.py``).

Order contract: ``for i in a, j in b, k in c`` in Julia nests with the LAST range innermost, which is
numpy's C-order ravel of an ``indexing='ij'`` mesh.  Points are produced in that order.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable

import numpy as np

__all__ = [
    "Box", "Rectangle", "Circle", "Ball", "BooleanUnion", "BooleanIntersection", "BooleanDifference",
    "Specification", "HalfSpace", "BoundaryLayer", "Grid", "Squaregrid", "Hexagrid", "CubicGrid", "covering",
    "generate_positions", "boundarybox",
    "Ellipse", "Ellipsoid", "Transform", "Polygon", "ClosedSpline", "Cone", "RevolutionBody",
    "VogelGrid", "BodycenteredGrid", "FacecenteredGrid", "DiamondGrid", "make_grid",
]


class Shape:
    """Supertype for shapes (src/structs.jl:19).  ``is_inside`` takes an (n,3) array."""

    def is_inside(self, X: np.ndarray) -> np.ndarray:  # pragma: no cover - abstract
        raise NotImplementedError

    def boundarybox(self) -> "Box":  # pragma: no cover - abstract
        raise NotImplementedError

    # src/geometry.jl:237-239
    def __add__(self, other: "Shape") -> "Shape":
        return BooleanUnion(self, other)

    def __sub__(self, other: "Shape") -> "Shape":
        return BooleanDifference(self, other)

    def __mul__(self, other: "Shape") -> "Shape":
        return BooleanIntersection(self, other)


@dataclass(frozen=True)
class Box(Shape):
    """src/geometry.jl:15-34 — closed on every face."""
    x1_min: float
    x2_min: float
    x3_min: float
    x1_max: float
    x2_max: float
    x3_max: float

    def is_inside(self, X):
        return ((self.x1_min <= X[:, 0]) & (X[:, 0] <= self.x1_max) & (self.x2_min <= X[:, 1])
                & (X[:, 1] <= self.x2_max) & (self.x3_min <= X[:, 2]) & (X[:, 2] <= self.x3_max))

    def boundarybox(self):
        return self

    @property
    def lo(self):
        return (self.x1_min, self.x2_min, self.x3_min)

    @property
    def hi(self):
        return (self.x1_max, self.x2_max, self.x3_max)


def Rectangle(x1_min, x2_min, x1_max, x2_max) -> Box:
    """src/geometry.jl:41-43"""
    return Box(float(x1_min), float(x2_min), 0.0, float(x1_max), float(x2_max), 0.0)


@dataclass(frozen=True)
class Circle(Shape):
    """src/geometry.jl:49-68"""
    x1: float
    x2: float
    r: float

    def is_inside(self, X):
        a = X[:, 0] - self.x1
        b = X[:, 1] - self.x2
        return a * a + b * b <= self.r * self.r

    def boundarybox(self):
        return Rectangle(self.x1 - self.r, self.x2 - self.r, self.x1 + self.r, self.x2 + self.r)


@dataclass(frozen=True)
class Ball(Shape):
    """src/geometry.jl:245-258"""
    x1: float
    x2: float
    x3: float
    r: float

    def is_inside(self, X):
        a = X[:, 0] - self.x1
        b = X[:, 1] - self.x2
        c = X[:, 2] - self.x3
        return a * a + b * b + c * c <= self.r * self.r

    def boundarybox(self):
        return Box(self.x1 - self.r, self.x2 - self.r, self.x3 - self.r, self.x1 + self.r, self.x2 + self.r,
                   self.x3 + self.r)


@dataclass(frozen=True)
class BooleanUnion(Shape):
    """src/geometry.jl:108-127"""
    s1: Shape
    s2: Shape

    def is_inside(self, X):
        return self.s1.is_inside(X) | self.s2.is_inside(X)

    def boundarybox(self):
        r1, r2 = self.s1.boundarybox(), self.s2.boundarybox()
        return Box(min(r1.x1_min, r2.x1_min), min(r1.x2_min, r2.x2_min), min(r1.x3_min, r2.x3_min),
                   max(r1.x1_max, r2.x1_max), max(r1.x2_max, r2.x2_max), max(r1.x3_max, r2.x3_max))


@dataclass(frozen=True)
class BooleanIntersection(Shape):
    """src/geometry.jl:134-153"""
    s1: Shape
    s2: Shape

    def is_inside(self, X):
        return self.s1.is_inside(X) & self.s2.is_inside(X)

    def boundarybox(self):
        r1, r2 = self.s1.boundarybox(), self.s2.boundarybox()
        return Box(max(r1.x1_min, r2.x1_min), max(r1.x2_min, r2.x2_min), max(r1.x3_min, r2.x3_min),
                   min(r1.x1_max, r2.x1_max), min(r1.x2_max, r2.x2_max), min(r1.x3_max, r2.x3_max))


@dataclass(frozen=True)
class BooleanDifference(Shape):
    """src/geometry.jl:160-171"""
    s1: Shape
    s2: Shape

    def is_inside(self, X):
        return self.s1.is_inside(X) & ~self.s2.is_inside(X)

    def boundarybox(self):
        return self.s1.boundarybox()


@dataclass(frozen=True)
class Specification(Shape):
    """src/geometry.jl:178-189.  ``f`` maps an (n,3) array to a boolean mask."""
    s: Shape
    f: Callable[[np.ndarray], np.ndarray]

    def is_inside(self, X):
        return np.asarray(self.f(X), dtype=bool) & self.s.is_inside(X)

    def boundarybox(self):
        return self.s.boundarybox()


@dataclass(frozen=True)
class HalfSpace:
    """A predicate for ``Specification`` that can also run on the device: ``x[axis] op bound`` with op one of
    ``< <= > >=`` — the form of every Specification lambda in the reference's examples (e.g.
    ``x -> (x[2] < box_height)``, collapse3d.jl:74-75, cavity_flow.jl:62-63)."""
    axis: int
    op: str
    bound: float

    def __call__(self, X):
        v = X[:, self.axis]
        return {"<": v < self.bound, "<=": v <= self.bound, ">": v > self.bound, ">=": v >= self.bound}[self.op]


@dataclass(frozen=True)
class Ellipse(Shape):
    """src/geometry.jl:70-98"""
    x1: float
    x2: float
    r1: float
    r2: float

    def __post_init__(self):
        if self.r1 <= 0.0 or self.r2 <= 0.0:
            raise ValueError("Degenerate ellipse definition (r <= 0)!")  # the reference logs @error and carries on

    def is_inside(self, X):
        a = (X[:, 0] - self.x1) / self.r1
        b = (X[:, 1] - self.x2) / self.r2
        return a * a + b * b <= 1

    def boundarybox(self):
        return Rectangle(self.x1 - self.r1, self.x2 - self.r2, self.x1 + self.r1, self.x2 + self.r2)


@dataclass(frozen=True)
class Ellipsoid(Shape):
    """src/geometry.jl:262-282"""
    x1: float
    x2: float
    x3: float
    r1: float
    r2: float
    r3: float

    def is_inside(self, X):
        a = (X[:, 0] - self.x1) / self.r1
        b = (X[:, 1] - self.x2) / self.r2
        c = (X[:, 2] - self.x3) / self.r3
        return a * a + b * b + c * c <= 1.0

    def boundarybox(self):
        return Box(self.x1 - self.r1, self.x2 - self.r2, self.x3 - self.r3, self.x1 + self.r1, self.x2 + self.r2,
                   self.x3 + self.r3)


class Transform(Shape):
    """src/geometry.jl:284-316: the shape ``s`` under x -> A x + b (A a 3x3 matrix, b a 3-vector)."""

    def __init__(self, s: Shape, A=None, b=None):
        self.s = s
        self.A = np.eye(3) if A is None else np.asarray(A, dtype=np.float64).reshape(3, 3)
        self.A_inv = np.linalg.inv(self.A)
        self.b = np.zeros(3) if b is None else np.asarray(b, dtype=np.float64).reshape(3)

    def is_inside(self, X):
        return self.s.is_inside((X - self.b) @ self.A_inv.T)

    def boundarybox(self):
        box = self.s.boundarybox()
        corners = np.array([[x, y, z] for x in (box.x1_min, box.x1_max) for y in (box.x2_min, box.x2_max)
                            for z in (box.x3_min, box.x3_max)])
        Y = corners @ self.A.T + self.b
        lo, hi = Y.min(axis=0), Y.max(axis=0)
        return Box(float(lo[0]), float(lo[1]), float(lo[2]), float(hi[0]), float(hi[1]), float(hi[2]))


class Polygon(Shape):
    """src/geometry.jl:318-360: winding-number test, half-open in y as in the reference."""

    def __init__(self, *points):
        self.xs = np.array([float(p[0]) for p in points])
        self.ys = np.array([float(p[1]) for p in points])
        self.deg = len(points)

    def is_inside(self, X):
        x_, y_ = X[:, 0], X[:, 1]
        wn = np.zeros(len(X), dtype=np.int64)
        for i in range(self.deg):
            nxt = (i + 1) % self.deg
            isleft = (self.xs[nxt] - self.xs[i]) * (y_ - self.ys[i]) - (x_ - self.xs[i]) * (self.ys[nxt] - self.ys[i])
            wn += ((self.ys[i] <= y_) & (y_ < self.ys[nxt]) & (isleft > 0.0)).astype(np.int64)
            wn -= ((self.ys[i] > y_) & (y_ >= self.ys[nxt]) & (isleft < 0.0)).astype(np.int64)
        return wn != 0

    def boundarybox(self):
        return Rectangle(float(self.xs.min()), float(self.ys.min()), float(self.xs.max()), float(self.ys.max()))


def ClosedSpline(*points, n: int = 32) -> Polygon:
    """src/geometry.jl:362-373: the polygon through ``n`` samples of the natural cubic spline through the points
    (closed by repeating the first one).  Interpolations.jl's ``BSpline(Cubic(Natural(OnGrid())))`` is the natural
    interpolating cubic spline, which is what scipy's ``CubicSpline(bc_type="natural")`` evaluates."""
    from scipy.interpolate import CubicSpline
    xs = [float(p[0]) for p in points] + [float(points[0][0])]
    ys = [float(p[1]) for p in points] + [float(points[0][1])]
    ts = np.arange(len(points) + 1) / len(points)
    sx, sy = CubicSpline(ts, xs, bc_type="natural"), CubicSpline(ts, ys, bc_type="natural")
    fine = [i / (n - 1) for i in range(n)]
    return Polygon(*[(float(sx(t)), float(sy(t))) for t in fine])


class Cone(Shape):
    """src/geometry.jl:375-413, literally (the axial coordinate ``s`` is NOT normalised by the length there, so the
    shape is a cone only for |b - a| = 1, as in the reference's own test)."""

    def __init__(self, a1, a2, a3, b1, b2, b3, ar, br):
        self.a = np.array([a1, a2, a3], dtype=np.float64)
        self.b = np.array([b1, b2, b3], dtype=np.float64)
        self.ar, self.br = float(ar), float(br)
        d = self.a - self.b
        self.len = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])

    def is_inside(self, X):
        s = (X - self.a) @ (self.b - self.a)
        ok = (0.0 <= s) & (s <= self.len)
        Y = X - s[:, None] * self.b - (1 - s)[:, None] * self.a
        t = np.sqrt(Y[:, 0] * Y[:, 0] + Y[:, 1] * Y[:, 1] + Y[:, 2] * Y[:, 2])
        return ok & (s / self.len * self.br + (1.0 - s / self.len) * self.ar >= t)

    def boundarybox(self):
        R = max(self.ar, self.br)
        lo, hi = np.minimum(self.a, self.b) - R, np.maximum(self.a, self.b) + R
        return Box(float(lo[0]), float(lo[1]), float(lo[2]), float(hi[0]), float(hi[1]), float(hi[2]))


@dataclass(frozen=True)
class RevolutionBody(Shape):
    """src/geometry.jl:415-437: the 2-D shape ``s`` (r, z) rotated around the z axis."""
    s: Shape

    def is_inside(self, X):
        r = np.sqrt(X[:, 0] * X[:, 0] + X[:, 1] * X[:, 1])
        return self.s.is_inside(np.column_stack([r, X[:, 2], np.zeros(len(X))]))

    def boundarybox(self):
        rect = self.s.boundarybox()
        R = rect.x1_max
        return Box(-R, -R, rect.x2_min, R, R, rect.x2_max)


class Grid:
    dim = 0


@dataclass(frozen=True)
class Squaregrid(Grid):
    """src/grids.jl:48-66"""
    dr: float
    dim = 2


class Hexagrid(Grid):
    """src/grids.jl:68-91"""
    dim = 2

    def __init__(self, dr: float):
        self.dr = dr
        self.a = (4 / 3) ** (1 / 4) * dr
        self.b = (3 / 4) ** (1 / 4) * dr


@dataclass(frozen=True)
class CubicGrid(Grid):
    """src/grids.jl:124-144"""
    dr: float
    dim = 3


GOLDEN_ANGLE = 2.39996322972865332  # src/grids.jl:7


class VogelGrid(Grid):
    """src/grids.jl:93-122: Vogel's spiral around the origin, one point per area dr^2."""
    dim = 2

    def __init__(self, dr: float):
        self.dr = dr
        self.k = dr / math.sqrt(math.pi)
        self.center = np.zeros(3)


@dataclass(frozen=True)
class BodycenteredGrid(Grid):
    """src/grids.jl:146-175"""
    dr: float
    dim = 3


@dataclass(frozen=True)
class FacecenteredGrid(Grid):
    """src/grids.jl:177-212"""
    dr: float
    dim = 3


@dataclass(frozen=True)
class DiamondGrid(Grid):
    """src/grids.jl:214-239"""
    dr: float
    dim = 3


def make_grid(dr: float, symm: str) -> Grid:
    """Grid(dr, symm), src/grids.jl:27-38."""
    kinds = {"square": Squaregrid, "hexagonal": Hexagrid, "vogel": VogelGrid, "cubic": CubicGrid,
             "facecentered": FacecenteredGrid, "bodycentered": BodycenteredGrid, "diamond": DiamondGrid}
    if symm not in kinds:
        raise ValueError("Unsupported grid type: " + symm)
    return kinds[symm](dr)


def boundarybox(s: Shape) -> Box:
    return s.boundarybox()


_CHUNK = 4_000_000  # lattice points evaluated per batch


def _ifloor(x: float) -> int:
    return int(math.floor(x))


def _iceil(x: float) -> int:
    return int(math.ceil(x))


def covering(grid: Grid, s: Shape) -> np.ndarray:
    """All lattice points of ``grid`` inside ``s`` as an (n,3) Float64 array, in the reference's order."""
    box = s.boundarybox()
    out = []
    if isinstance(grid, Squaregrid):
        i0, j0 = _ifloor(box.x1_min / grid.dr), _ifloor(box.x2_min / grid.dr)
        i1, j1 = _iceil(box.x1_max / grid.dr), _iceil(box.x2_max / grid.dr)
        js = np.arange(j0, j1 + 1, dtype=np.int64)
        rows = max(1, _CHUNK // max(1, len(js)))
        for ia in range(i0, i1 + 1, rows):
            is_ = np.arange(ia, min(ia + rows, i1 + 1), dtype=np.int64)
            I, J = np.meshgrid(is_, js, indexing="ij")
            X = np.empty((I.size, 3))
            X[:, 0] = I.ravel() * grid.dr
            X[:, 1] = J.ravel() * grid.dr
            X[:, 2] = 0.0
            out.append(X[s.is_inside(X)])
    elif isinstance(grid, Hexagrid):
        i0, j0 = _ifloor(box.x1_min / grid.a) - 1, _ifloor(box.x2_min / grid.b)
        i1, j1 = _iceil(box.x1_max / grid.a), _iceil(box.x2_max / grid.b)
        js = np.arange(j0, j1 + 1, dtype=np.int64)
        rows = max(1, _CHUNK // max(1, len(js)))
        for ia in range(i0, i1 + 1, rows):
            is_ = np.arange(ia, min(ia + rows, i1 + 1), dtype=np.int64)
            I, J = np.meshgrid(is_, js, indexing="ij")
            I, J = I.ravel(), J.ravel()
            X = np.empty((I.size, 3))
            # j % 2 in Julia is the C remainder (sign of the dividend), src/grids.jl:83
            X[:, 0] = (I + np.fmod(J, 2) / 2) * grid.a
            X[:, 1] = J * grid.b
            X[:, 2] = 0.0
            out.append(X[s.is_inside(X)])
    elif isinstance(grid, CubicGrid):
        i0, j0, k0 = (_ifloor(box.x1_min / grid.dr), _ifloor(box.x2_min / grid.dr), _ifloor(box.x3_min / grid.dr))
        i1, j1, k1 = (_iceil(box.x1_max / grid.dr), _iceil(box.x2_max / grid.dr), _iceil(box.x3_max / grid.dr))
        js = np.arange(j0, j1 + 1, dtype=np.int64)
        ks = np.arange(k0, k1 + 1, dtype=np.int64)
        rows = max(1, _CHUNK // max(1, len(js) * len(ks)))
        for ia in range(i0, i1 + 1, rows):
            is_ = np.arange(ia, min(ia + rows, i1 + 1), dtype=np.int64)
            I, J, K = np.meshgrid(is_, js, ks, indexing="ij")
            X = np.empty((I.size, 3))
            X[:, 0] = I.ravel() * grid.dr
            X[:, 1] = J.ravel() * grid.dr
            X[:, 2] = K.ravel() * grid.dr
            out.append(X[s.is_inside(X)])
    elif isinstance(grid, VogelGrid):  # src/grids.jl:104-122
        corners = np.array([[box.x1_min, box.x2_min, 0.0], [box.x1_max, box.x2_min, 0.0], [box.x1_max, box.x2_max, 0.0],
                            [box.x1_min, box.x2_max, 0.0]]) - grid.center
        R = float(np.max(np.sqrt(corners[:, 0] ** 2 + corners[:, 1] ** 2 + corners[:, 2] ** 2)))
        N = (R / grid.k) ** 2
        for n0 in range(1, int(math.floor(N)) + 1, _CHUNK):
            n = np.arange(n0, min(n0 + _CHUNK, int(math.floor(N)) + 1), dtype=np.float64)  # for n in 1:N (N a Float64)
            rad = grid.k * np.sqrt(n)
            X = np.empty((n.size, 3))
            X[:, 0] = grid.center[0] + rad * np.cos(n * GOLDEN_ANGLE)
            X[:, 1] = grid.center[1] + rad * np.sin(n * GOLDEN_ANGLE)
            X[:, 2] = grid.center[2] + rad * 0.0
            out.append(X[s.is_inside(X)])
    elif isinstance(grid, (BodycenteredGrid, FacecenteredGrid)):  # src/grids.jl:150-175, 181-212
        a = (2 ** (1 / 3) if isinstance(grid, BodycenteredGrid) else 4 ** (1 / 3)) * grid.dr
        i0, j0, k0 = _ifloor(box.x1_min / a), _ifloor(box.x2_min / a), _ifloor(box.x3_min / a)
        i1, j1, k1 = _iceil(box.x1_max / a), _iceil(box.x2_max / a), _iceil(box.x3_max / a)
        js = np.arange(j0, j1 + 1, dtype=np.int64)
        ks = np.arange(k0, k1 + 1, dtype=np.int64)
        rows = max(1, _CHUNK // max(1, len(js) * len(ks)))
        # first loop: cell corners; second loop: the centred points, interleaved per (i, j, k) as the reference pushes them
        if isinstance(grid, BodycenteredGrid):
            passes = [[(0.0, 0.0, 0.0)], [(0.5, 0.5, 0.5)]]
        else:
            passes = [[(0.0, 0.0, 0.0)], [(0.5, 0.5, 0.0), (0.5, 0.0, 0.5), (0.0, 0.5, 0.5)]]
        for shifts in passes:
            for ia in range(i0, i1 + 1, rows):
                is_ = np.arange(ia, min(ia + rows, i1 + 1), dtype=np.int64)
                I, J, K = (m.ravel() for m in np.meshgrid(is_, js, ks, indexing="ij"))
                X = np.empty((I.size, len(shifts), 3))
                for t, (si, sj, sk) in enumerate(shifts):
                    X[:, t, 0] = (I + si) * a
                    X[:, t, 1] = (J + sj) * a
                    X[:, t, 2] = (K + sk) * a
                X = X.reshape(-1, 3)
                out.append(X[s.is_inside(X)])
    elif isinstance(grid, DiamondGrid):  # src/grids.jl:218-239
        a = 0.5 * grid.dr
        i0, j0, k0 = _ifloor(box.x1_min / a), _ifloor(box.x2_min / a), _ifloor(box.x3_min / a)
        i1, j1, k1 = _iceil(box.x1_max / a), _iceil(box.x2_max / a), _iceil(box.x3_max / a)
        js_all = np.arange(j0, j1 + 1, dtype=np.int64)
        ks_all = np.arange(k0, k1 + 1, dtype=np.int64)
        for i in range(i0, i1 + 1):
            par = i & 1  # isodd(i) == isodd(j) == isodd(k)
            J, K = (m.ravel() for m in np.meshgrid(js_all[(js_all & 1) == par], ks_all[(ks_all & 1) == par], indexing="ij"))
            keep = np.mod(i + J + K, 4) <= 1  # ((i+j+k) % 4 + 4) % 4 in {0, 1}
            J, K = J[keep], K[keep]
            X = np.empty((J.size, 3))
            X[:, 0] = i * a
            X[:, 1] = J * a
            X[:, 2] = K * a
            out.append(X[s.is_inside(X)])
    else:
        raise TypeError("unsupported grid")
    return np.concatenate(out, axis=0) if out else np.zeros((0, 3))


class BoundaryLayer(Shape):
    """src/geometry.jl:198-234: not in ``s`` but ``x + dx`` in ``s`` for a lattice offset |dx| <= width."""

    def __init__(self, s: Shape, grid: Grid, width: float):
        self.s = s
        self.dim = grid.dim
        self.dxs = covering(grid, Ball(0.0, 0.0, 0.0, width))
        self.width = width

    def is_inside(self, X):
        res = np.zeros(len(X), dtype=bool)
        cand = np.flatnonzero(~self.s.is_inside(X))
        if cand.size:
            Xc = X[cand]
            hit = np.zeros(cand.size, dtype=bool)
            for dx in self.dxs:
                todo = np.flatnonzero(~hit)
                if todo.size == 0:
                    break
                hit[todo] = self.s.is_inside(Xc[todo] + dx)
            res[cand] = hit
        return res

    def boundarybox(self):
        r = self.s.boundarybox()
        w = self.width
        if self.dim == 2:
            return Rectangle(r.x1_min - w, r.x2_min - w, r.x1_max + w, r.x2_max + w)
        return Box(r.x1_min - w, r.x2_min - w, r.x3_min - w, r.x1_max + w, r.x2_max + w, r.x3_max + w)


def generate_positions(grid: Grid, geometry: Shape) -> np.ndarray:
    """Positions ``generate_particles!`` would push, in order (src/grids.jl:253-258)."""
    return covering(grid, geometry)


# --------------------------------------------------------------------------- device generation (sp_generate.cu)
def lattice_index_box(grid: Grid, s: Shape):
    """(i0, i1, j0, j1, k0, k1), inclusive: the index ranges ``covering`` loops over (src/grids.jl:53-56, 76-79,
    129-134)."""
    box = s.boundarybox()
    if isinstance(grid, Squaregrid):
        return (_ifloor(box.x1_min / grid.dr), _iceil(box.x1_max / grid.dr), _ifloor(box.x2_min / grid.dr),
                _iceil(box.x2_max / grid.dr), 0, 0)
    if isinstance(grid, Hexagrid):
        return (_ifloor(box.x1_min / grid.a) - 1, _iceil(box.x1_max / grid.a), _ifloor(box.x2_min / grid.b),
                _iceil(box.x2_max / grid.b), 0, 0)
    if isinstance(grid, CubicGrid):
        return (_ifloor(box.x1_min / grid.dr), _iceil(box.x1_max / grid.dr), _ifloor(box.x2_min / grid.dr),
                _iceil(box.x2_max / grid.dr), _ifloor(box.x3_min / grid.dr), _iceil(box.x3_max / grid.dr))
    raise TypeError("unsupported grid")


def compile_shape(s: Shape):
    """Postfix program of the shape for ``sp_generate_particles``: a list of (kind, a, b, params) with children
    before parents and the root last, plus the lattice offsets of the (single) BoundaryLayer if there is one."""
    from . import abi
    K = abi.K
    nodes, offsets = [], [None]

    def emit(kind, a=0, b=0, p=()):
        nodes.append((K[kind], a, b, tuple(float(v) for v in p)))
        return len(nodes) - 1

    def walk(sh):
        if isinstance(sh, Box):
            return emit("SP_SHAPE_BOX", p=sh.lo + sh.hi)
        if isinstance(sh, Circle):
            return emit("SP_SHAPE_CIRCLE", p=(sh.x1, sh.x2, sh.r * sh.r))
        if isinstance(sh, Ball):
            return emit("SP_SHAPE_BALL", p=(sh.x1, sh.x2, sh.x3, sh.r * sh.r))
        if isinstance(sh, (BooleanUnion, BooleanIntersection, BooleanDifference)):
            a, b = walk(sh.s1), walk(sh.s2)
            kind = {BooleanUnion: "SP_SHAPE_UNION", BooleanIntersection: "SP_SHAPE_INTERSECTION",
                    BooleanDifference: "SP_SHAPE_DIFFERENCE"}[type(sh)]
            return emit(kind, a, b)
        if isinstance(sh, Specification):
            if not isinstance(sh.f, HalfSpace):
                raise TypeError("only HalfSpace predicates can run on the device (a Python lambda cannot)")
            a = walk(sh.s)
            b = emit("SP_SHAPE_HALFSPACE", sh.f.axis, {"<": 0, "<=": 1, ">": 2, ">=": 3}[sh.f.op], (sh.f.bound,))
            return emit("SP_SHAPE_INTERSECTION", b, a)   # f(x) && is_inside(x, s), geometry.jl:183-185
        if isinstance(sh, BoundaryLayer):
            if offsets[0] is not None:
                raise TypeError("one BoundaryLayer per shape on the device")
            a = walk(sh.s)
            offsets[0] = np.ascontiguousarray(sh.dxs, dtype=np.float64)
            return emit("SP_SHAPE_BOUNDARY_LAYER", a)
        raise TypeError(f"shape {type(sh).__name__} is not supported by the device generator")

    walk(s)
    return nodes, offsets[0]
