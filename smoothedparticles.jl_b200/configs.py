"""The BASELINE configs as host programs over registered operators.

Each ``Case`` restates one reference script: constants, particle struct, initial state
(``make_system``) and the body of its time loop, calling ``apply`` / ``create_cell_list`` in the same
order as the script.  A Case is backend-agnostic: ``case.make(ParticleSystem)`` builds it on the GPU,
``case.make(OracleSystem)`` on the CPU oracle (tests only), and ``case.step(sys)`` advances either.

  collapse_dry           examples/collapse_dry.jl            2-D WCSPH dam break
  collapse3d             examples/collapse3d.jl              3-D WCSPH dam break (dr scalable to 10 M)
  cavity_flow            examples/cavity_flow.jl             2-D lid-driven cavity
  collapse_dry_implicit  examples/collapse_dry_implicit.jl   2-D ISPH, matrix-free pressure Poisson + CG
  collision_2d           tests/test_collision_2d.jl          two colliding discs (the reference's own test)
  static_container       examples/static_container.jl        tank at rest, density integrated in the pair loop
  drop                   examples/drop.jl                    3-D drop with colour-field surface tension
  collapse_symplectic    examples/collapse_symplectic.jl     reversible fixed-point Verlet, Lennard-Jones walls
  kepler_vortex          examples/Kepler_vortex.jl           fluid ring in central gravity, same integrator
  cylinder               examples/cylinder.jl                channel flow past a cylinder, inflow buffer, per-particle mass
  rod                    examples/rod.jl                     elastic rod, tensor-valued particle fields
  shtc_ldc               examples/SHTC/ldc.jl                lid-driven cavity with the SHTC model (3x3 distortion field)
  shtc_beryllium         examples/SHTC/beryllium.jl          vibrating beryllium plate, SHTC solid (2-D)
  shtc_twist3d           examples/SHTC/twist3d.jl            twisting rubber column, SHTC solid (3-D)
  shtc_taco              examples/SHTC/taco.jl               Taylor-Couette flow, SHTC fluid on a Vogel spiral
  lattice_box            synthetic S1 block of SURVEY §8(d)  jittered cubic lattice, all fluid
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

from . import abi, geometry as geo, operators as ops

K = abi.K


@dataclass
class Case:
    name: str
    fields: Dict[str, int]
    domain: geo.Box
    h: float
    init: Dict[str, np.ndarray]          # initial field arrays in reference order ("x" required)
    step: Callable                        # step(sys): one pass of the script's time loop body
    prologue: Callable = lambda sys: None  # what the script does once before the loop
    consts: Dict[str, float] = field(default_factory=dict)
    program: Optional[int] = None         # SP_PROGRAM_* equivalent of step(), if any
    program_fields: tuple = ()
    program_params: tuple = ()
    dim: int = 3
    recipe: tuple = ()                    # ((grid, shape, {field: constant}), ...): the generate_particles! calls

    @property
    def n(self) -> int:
        return len(self.init["x"])

    def make(self, system_cls, **kw):
        sys = system_cls(self.fields, self.domain, self.h, **kw)
        sys.add_particles(**self.init)
        return sys

    def make_on_device(self, system_cls, **kw):
        """The same initial state built by the device generator (sp_generate_particles) instead of a host upload:
        one generate_particles call per entry of ``recipe``, in the script's order."""
        if not self.recipe:
            raise ValueError(f"{self.name}: no device recipe")
        sys = system_cls(self.fields, self.domain, self.h, **kw)
        for grid, shape, constants in self.recipe:
            sys.generate_particles(grid, shape, **constants)
        return sys


# --------------------------------------------------------------------------- collapse_dry.jl
def collapse_dry(dr: float = 1.5e-2) -> Case:
    """examples/collapse_dry.jl:44-102 (constants, make_system) and :194-211 (loop)."""
    h = 3.0 * dr
    rho0 = 1000.0
    m = rho0 * dr ** 2
    c = 50.0
    g = (0.0, -7.0, 0.0)  # -7.0*VECY
    mu = 8.4e-4
    nu = 1.0e-6
    wcw, wch, bh, bw = 1.0, 2.0, 3.0, 4.0
    wall_width = 2.5 * dr
    dt = 0.1 * h / c
    grid = geo.Hexagrid(dr)
    box = geo.Rectangle(0.0, 0.0, bw, bh)
    fluid = geo.Rectangle(0.0, 0.0, wcw, wch)
    walls = geo.BoundaryLayer(box, grid, wall_width)
    walls = geo.Specification(walls, lambda X: X[:, 1] < bh)
    domain = (box + walls).boundarybox()
    xf = geo.covering(grid, fluid)
    xw = geo.covering(grid, walls)
    x = np.concatenate([xf, xw])
    typ = np.concatenate([np.zeros(len(xf)), np.ones(len(xw))])
    P = rho0 * g[1] * (x[:, 1] - wch)       # :98 hydrostatic pressure
    rho = rho0 + P / c ** 2                 # :99
    fields = {"v": 3, "Dv": 3, "rho": 1, "Drho": 1, "P": 1, "type": 1}
    init = {"x": x, "rho": rho, "P": P, "type": typ}
    o_bom = ops.balance_of_mass("wendland2", m, h, nu)
    o_fp = ops.find_pressure(dt, c, rho0)
    o_if = ops.internal_force("wendland2", m, h, mu, rho0)
    o_mv = ops.move(0.5 * dt)
    o_ac = ops.accelerate(0.5 * dt, g)

    def prologue(sys):  # :200-201
        sys.create_cell_list()
        sys.apply(o_if)

    def step(sys):  # :203-211
        sys.apply(o_ac)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_bom)
        sys.apply(o_fp)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_if)
        sys.apply(o_ac)

    return Case("collapse_dry", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, rho0=rho0, m=m, c=c, mu=mu, nu=nu, dt=dt, g=g, width=wcw, height=wch),
                program=K["SP_PROGRAM_WCSPH_2D"], program_fields=("x", "v", "Dv", "rho", "Drho", "P", "type"),
                program_params=(float(K["SP_KERNEL_WENDLAND2"]), m, h, 2 * nu, dt, c * c, rho0, mu, *g), dim=2)


# --------------------------------------------------------------------------- collapse3d.jl
def relabel_axes(axis: int):
    """Cyclic relabelling of the coordinate axes that makes physical axis ``axis`` the third one: column k of a relabelled
    vector is column ``perm[k]`` of the physical one (cyclic, so handedness is kept).  The slab decomposition cuts along the
    slowest axis of the cell key (z in 3-D, csrc/sp_slab.cu); a host that wants to cut along x or y hands the library
    relabelled coordinates — every registered pair operator depends on positions through differences and distances only,
    and vector parameters (gravity) are relabelled with them — and relabels vector fields back after a download
    (``inverse``: physical[:, perm] = relabelled)."""
    assert axis in (0, 1, 2)
    return [(axis + 1) % 3, (axis + 2) % 3, axis]


def collapse3d(dr: float = 5.0e-3, depth_scale: float = 1.0, z_range=None, slab_axis: int = 2) -> Case:
    """examples/collapse3d.jl:28-84 and :134-151.

    ``slab_axis`` (0, 1 or 2): the physical axis a slab decomposition should cut along.  2 (default) is the script as it
    stands; 0 / 1 hand the library cyclically relabelled coordinates (``relabel_axes``) so that axis becomes the slowest
    axis of the cell key — same physics, same neighbour sets, sums in a different visiting order (``case.consts["perm"]``
    maps back).  The device recipe is dropped for a relabelled case (host generation + upload).

    ``depth_scale`` extrudes the box (and the water column) along z — the direction the dam break is invariant
    in — for weak scaling over slabs; ``z_range=(lo, hi)`` generates only the lattice points with lo <= z < hi
    (one rank's share), the domain stays the global one.

    The shipped script does not run (``import`` instead of ``using``; ``rho`` undefined in
    internal_force!, :101).  Adopted correction: the pressure term of
    collapse_dry.jl:138, ``p.P/p.rho^2 + q.P/q.rho^2``, with rDwendland3.
    """
    h = 2.0 * dr
    rho0 = 1000.0
    m = rho0 * dr ** 3
    c = 50.0
    g = (0.0, 0.0, -9.8)  # -9.8*VECZ
    mu = 8.4e-4
    nu = 1.0e-4
    wcw, wch, bh, bw, bd = 0.142, 0.293, 0.35, 0.584, 0.15 * depth_scale
    wall_width = 2.5 * dr
    dt = 0.1 * h / c
    grid = geo.CubicGrid(dr)
    box = geo.Box(0.0, 0.0, 0.0, bw, bh, bd)
    fluid = geo.Box(0.0, 0.0, 0.0, wcw, wch, bd)
    walls = geo.BoundaryLayer(box, grid, wall_width)
    walls = geo.Specification(walls, geo.HalfSpace(1, "<", bh))   # x -> (x[2] < box_height), :74-75
    domain = walls.boundarybox()
    if z_range is not None:
        zlo, zhi = z_range
        big = 1e30
        zslab = geo.Box(-big, -big, zlo, big, big, zhi)
        fluid_g = geo.Specification(fluid * zslab, geo.HalfSpace(2, "<", zhi))
        walls_g = geo.Specification(walls * zslab, geo.HalfSpace(2, "<", zhi))
    else:
        fluid_g, walls_g = fluid, walls
    xf = geo.covering(grid, fluid_g)
    xw = geo.covering(grid, walls_g)
    x = np.concatenate([xf, xw])
    typ = np.concatenate([np.zeros(len(xf)), np.ones(len(xw))])
    fields = {"v": 3, "Dv": 3, "P": 1, "rho": 1, "Drho": 1, "type": 1}
    init = {"x": x, "rho": np.full(len(x), rho0), "type": typ}
    o_bom = ops.balance_of_mass("wendland3", m, h, nu)
    o_fp = ops.find_pressure(dt, c, rho0)
    o_if = ops.internal_force("wendland3", m, h, mu, rho0)
    o_mv = ops.move(dt)
    perm = relabel_axes(slab_axis)
    recipe = ((grid, fluid_g, {"rho": rho0, "type": 0.0}), (grid, walls_g, {"rho": rho0, "type": 1.0}))
    if slab_axis != 2:
        init["x"] = np.ascontiguousarray(x[:, perm])
        g = tuple(g[k] for k in perm)
        domain = geo.Box(*(domain.lo[k] for k in perm), *(domain.hi[k] for k in perm))
        recipe = ()
    o_ac = ops.accelerate(0.5 * dt, g)

    def step(sys):  # :136-150
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_bom)
        sys.apply(o_fp)
        sys.apply(o_if)
        sys.apply(o_ac)
        sys.apply(o_ac)

    return Case("collapse3d", fields, domain, h, init, step,
                consts=dict(dr=dr, h=h, rho0=rho0, m=m, c=c, mu=mu, nu=nu, dt=dt, g=g, perm=perm),
                program=K["SP_PROGRAM_WCSPH_3D"], program_fields=("x", "v", "Dv", "rho", "Drho", "P", "type"),
                program_params=(float(K["SP_KERNEL_WENDLAND3"]), m, h, 2 * nu, dt, c * c, rho0, mu, *g), dim=3, recipe=recipe)


def collapse3d_dr_for(n_target: float) -> float:
    """dr that scales examples/collapse3d.jl to about n_target particles (walls scale with area)."""
    # N(dr) ~ V_fluid/dr^3 + A_wall*2.5/dr^2 ; solve by bisection on the analytic estimate
    vf = 0.142 * 0.293 * 0.15
    aw = 2 * (0.584 * 0.35 + 0.35 * 0.15) + 0.584 * 0.15  # 4 side walls + floor, no lid
    lo, hi = 1e-4, 5e-2
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        n = vf / mid ** 3 + aw * 3.0 / mid ** 2
        if n > n_target:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


# --------------------------------------------------------------------------- cavity_flow.jl
def cavity_flow(N: int = 100, Re: int = 100) -> Case:
    """examples/cavity_flow.jl:28-86 (constants, make_system) and :137-151 (loop)."""
    llid = 1.0
    rho0 = 1.0
    vlid = 1.0
    dr = llid / N
    h = 3.0 * dr
    m = rho0 * dr ** 2
    c = 20 * vlid
    P0 = 5.0
    wwall = h
    dt = 0.1 * h / c
    grid = geo.Hexagrid(dr)
    box = geo.Rectangle(0.0, 0.0, llid, llid)
    wall = geo.BoundaryLayer(box, grid, wwall)
    domain = (box + wall).boundarybox()
    lid = geo.Specification(wall, lambda X: X[:, 1] > llid)
    wall2 = geo.Specification(wall, lambda X: X[:, 1] <= llid)
    xf, xl, xw = geo.covering(grid, box), geo.covering(grid, lid), geo.covering(grid, wall2)
    x = np.concatenate([xf, xl, xw])
    typ = np.concatenate([np.zeros(len(xf)), np.full(len(xl), 2.0), np.ones(len(xw))])
    fields = {"v": 3, "Dv": 3, "rho": 1, "Drho": 1, "P": 1, "type": 1}
    init = {"x": x, "rho": np.full(len(x), rho0), "type": typ}
    o_bom = ops.balance_of_mass("wendland2", m, h, 0.0)
    o_fp = ops.find_pressure(dt, c, rho0, P0)
    o_if = ops.internal_force_cavity(m, h, Re, vlid, ylid=1.0, lid_type=2.0)
    o_mv = ops.move(0.5 * dt)
    o_ac = ops.accelerate(0.5 * dt)

    def prologue(sys):  # :82-84
        sys.create_cell_list()
        sys.apply(o_fp)
        sys.apply(o_if)

    def step(sys):  # :138-150
        sys.apply(o_ac)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_bom)
        sys.apply(o_fp)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_if)
        sys.apply(o_ac)

    return Case("cavity_flow", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, rho0=rho0, m=m, c=c, dt=dt, Re=Re, P0=P0), dim=2)


# --------------------------------------------------------------------------- static_container.jl
def static_container(dr: float = 1.5e-3) -> Case:
    """examples/static_container.jl:25-45 (constants), :82-98 (make_system), :133-141 (loop): a tank at rest with
    the hydrostatic density profile; the density is integrated inside the pair loop and the pressure comes from the
    equation of state inside internal_force!."""
    h = 1.8 * dr
    rho0 = 1000.0
    m = rho0 * dr ** 2
    c = 40.0
    g = (0.0, -1.0, 0.0)  # -VECY
    mu = 8.4e-4
    water_depth, box_height, box_width = 0.14, 0.18, 0.14
    wall_width = 2.5 * dr
    dt = 0.2 * h / c
    grid = geo.Squaregrid(dr)
    box = geo.Rectangle(0.0, 0.0, box_width, box_height)
    fluid = geo.Rectangle(0.0, 0.0, box_width, water_depth)
    walls = geo.BoundaryLayer(box, grid, wall_width)
    domain = (box + walls).boundarybox()
    xf, xw = geo.covering(grid, fluid), geo.covering(grid, walls)
    x = np.concatenate([xf, xw])
    typ = np.concatenate([np.zeros(len(xf)), np.ones(len(xw))])
    P = rho0 * g[1] * (x[:, 1] - water_depth)          # hydrostatic pressure, :91
    rho = rho0 + P / c ** 2                            # :92
    fields = {"v": 3, "a": 3, "rho": 1, "type": 1}
    init = {"x": x, "rho": rho, "type": typ}
    o_bom = ops.sc_balance_of_mass("wendland2", m, h, dt)
    o_if = ops.sc_internal_force("wendland2", m, h, mu, c, rho0)
    o_mv = ops.move_all(0.5 * dt)
    o_ac = ops.accelerate(0.5 * dt, g, Dv="a")

    def prologue(sys):  # :94-95
        sys.create_cell_list()
        sys.apply(o_if)

    def step(sys):  # :133-141
        sys.apply(o_ac)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_bom)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_if)
        sys.apply(o_ac)

    return Case("static_container", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, rho0=rho0, m=m, c=c, mu=mu, dt=dt, g=g, water_depth=water_depth), dim=2)


# --------------------------------------------------------------------------- drop.jl
def drop(dr: float = 3.7e-5) -> Case:
    """examples/drop.jl:18-71 (constants, make_system), :139-152 (verlet_step!), :168-175 (initialisation): a water
    drop on a desk with colour-field surface tension.  h = 3 dr in 3-D: ~113 neighbours per particle."""
    h = 3.0 * dr
    rad = 1e-3
    deskw = 0.9 * h
    rho0 = 1000.0
    m = rho0 * dr ** 3
    mu = 0.1
    beta = 72e-3
    vol = dr ** 3
    g = (0.0, 0.0, -9.8)
    c = 10.0 * max(np.sqrt(beta / rho0 / dr), np.sqrt(4 * 9.8 * rad))
    dt = 0.3 * dr / c
    s0 = dr * dr / 100
    grid = geo.CubicGrid(dr)
    ball = geo.Ball(0.0, 0.0, rad + h, rad)
    desk = geo.Box(-2 * rad, -2 * rad, -deskw, 2 * rad, 2 * rad, 0.0)
    domain = geo.Box(-2 * rad, -2 * rad, -2 * deskw, 2 * rad, 2 * rad, 2.2 * rad)
    xf, xs = geo.covering(grid, ball), geo.covering(grid, desk)
    x = np.concatenate([xf, xs])
    typ = np.concatenate([np.zeros(len(xf)), np.ones(len(xs))])
    fields = {"v": 3, "a": 3, "P": 1, "rho": 1, "rho0": 1, "n": 3, "type": 1}
    init = {"x": x, "type": typ}
    o_rho = ops.density_sum("wendland3", m, h, out="rho")
    o_rho0 = ops.density_sum("wendland3", m, h, out="rho0")
    o_p = ops.pressure_from_rho(c)
    o_n = ops.find_normal("wendland3", vol, h)
    o_nn = ops.normalize(s0)
    o_f = ops.internal_force_tension(m, h, mu, rho0, beta, s0)
    o_ra, o_rr, o_rn = ops.fill("a", 0.0), ops.fill("rho", 0.0), ops.fill("n", 0.0)
    o_mv = ops.advect(dt)                     # x += (type==FLUID)*dt*v: the desk never gains a velocity
    o_ac = ops.accelerate(0.5 * dt, g, Dv="a")

    def prologue(sys):  # :168-175
        sys.create_cell_list()
        sys.apply(o_rho0, self_=True)
        sys.apply(o_rho, self_=True)
        sys.apply(o_p)
        sys.apply(o_n)
        sys.apply(o_nn)
        sys.apply(o_f)

    def step(sys):  # :139-152
        sys.apply(o_ac)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_rr)
        sys.apply(o_rho, self_=True)
        sys.apply(o_rn)
        sys.apply(o_n, self_=True)
        sys.apply(o_nn)
        sys.apply(o_p)
        sys.apply(o_ra)
        sys.apply(o_f)
        sys.apply(o_ac)

    return Case("drop", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, rho0=rho0, m=m, c=c, mu=mu, beta=beta, dt=dt, g=g, s0=s0, vol=vol), dim=3,
                recipe=((grid, ball, {"type": 0.0}), (grid, desk, {"type": 1.0})))


# --------------------------------------------------------------------------- collapse_symplectic.jl
def collapse_symplectic(dr: float = 1.0e-2) -> Case:
    """examples/collapse_symplectic.jl:39-95 (constants, make_system), :171-181 (verlet_step!), :198-203 (init):
    the dam break integrated with the reversible fixed-point Verlet scheme of utils/FixPA.jl; walls repel through
    a Lennard-Jones force instead of carrying pressure."""
    h = 3.0 * dr
    rho0 = 1000.0
    m = rho0 * dr ** 2
    g = (0.0, -9.8, 0.0)  # -9.8*VECY
    wcw, wch, bh, bw = 1.0, 2.0, 3.0, 4.0
    wall_width = 2.5 * dr
    c = 50.0
    dr_wall = 0.95 * dr
    E_wall = 10 * math.sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) * wch  # 10*norm(g)*water_column_height
    eps = 1e-16
    dt = 0.1 * h / c
    grid = geo.Squaregrid(dr)
    box = geo.Rectangle(0.0, 0.0, bw, bh)
    fluid = geo.Rectangle(0.0, 0.0, wcw, wch)
    walls = geo.BoundaryLayer(box, grid, wall_width)
    domain = geo.Rectangle(-bw, -bw, 2 * bw, 3 * bh)
    xf, xw = geo.covering(grid, fluid), geo.covering(grid, walls)
    x = np.concatenate([xf, xw])
    typ = np.concatenate([np.zeros(len(xf)), np.ones(len(xw))])
    fields = {"v": 3, "a": 3, "P": 1, "rho": 1, "rho0": 1, "type": 1, "U": 1}
    init = {"x": x, "type": typ}
    o_rho = ops.density_sum_fluid("wendland2", m, h, out="rho")
    o_rho0 = ops.density_sum_fluid("wendland2", m, h, out="rho0")
    o_p = ops.pressure_from_rho(c)
    o_f = ops.internal_force_lj("wendland2", m, h, dr_wall, E_wall, eps)
    o_ra = ops.fill("a", 0.0)
    o_rr = ops.fill("rho", 0.0)
    o_mv = ops.move_rev(dt)
    o_ac = ops.accelerate_rev(0.5 * dt, g)

    def prologue(sys):  # :198-203
        sys.create_cell_list()
        sys.apply(o_rho0, self_=True)
        sys.apply(o_rho, self_=True)
        sys.apply(o_p)
        sys.apply(o_f)

    def step(sys):  # verlet_step! :171-181
        sys.apply(o_ac)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_rr)
        sys.apply(o_rho, self_=True)
        sys.apply(o_p)
        sys.apply(o_ra)
        sys.apply(o_f)
        sys.apply(o_ac)

    return Case("collapse_symplectic", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, rho0=rho0, m=m, c=c, dt=dt, g=g, dr_wall=dr_wall, E_wall=E_wall, eps=eps),
                dim=2, recipe=((grid, fluid, {"type": 0.0}), (grid, walls, {"type": 1.0})))


# --------------------------------------------------------------------------- Kepler_vortex.jl
def kepler_ring_radii(N_rings: int = 25, r0: float = 10.0):
    """Kepler_vortex.jl:43-60,68: radii of the Gaussian rings, r_f(u) = inverse of the normalised cumulative
    surface density f(r) = int_0^r 2 pi s exp(-30 (1 - s/r0)^2) ds / int_0^40 (...).  The script tabulates f at
    0:0.5:25 with QuadGK (rtol 1e-3), interpolates the table with a cubic B-spline and inverts it with Roots.jl;
    here: scipy quad, a natural cubic spline through the same table and brentq.  Set-up only (no Julia output to
    compare with); the operators do not depend on it."""
    from scipy.integrate import quad
    from scipy.interpolate import CubicSpline
    from scipy.optimize import brentq

    def sigma(r):
        return 2 * math.pi * r * math.exp(-30 * (1 - r / r0) ** 2)

    den = quad(sigma, 0, 40, epsrel=1e-6)[0]
    rs = np.arange(0.0, 25.0 + 1e-9, 0.5)
    table = np.array([quad(sigma, 0, r, epsrel=1e-3)[0] / den for r in rs])
    spline = CubicSpline(rs, table, bc_type="natural")

    def r_f(F):
        return brentq(lambda r: float(spline(r)) - F, 2.0, 20.0, xtol=1e-14)

    us = 0.01 + (0.99 - 0.01) / N_rings * np.arange(N_rings + 1)  # 0.01:(0.99-0.01)/N_rings:0.99
    return np.array([r_f(u) for u in us]), r_f(0.25 + 1 / N_rings) - r_f(0.25)


def kepler_vortex(N_rings: int = 25) -> Case:
    """examples/Kepler_vortex.jl:28-99 (constants), :109-134 (rings of particles on Keplerian orbits),
    :220-230 (verlet_step!), :253-258 (init): a self-gravitating-free fluid ring around a central mass, integrated
    with the reversible fixed-point scheme; no wall particles are generated (the LJ branch stays idle)."""
    r0, GM = 10.0, 1000.0
    radii, dr = kepler_ring_radii(N_rings, r0)
    h = 3.0 * dr
    rho0 = 1.0
    m = rho0 * dr ** 2
    bw = 4 * r0
    c = 0.01
    dr_wall = 0.95 * dr
    E_wall = GM / r0
    eps = 1e-16
    dt = 0.0001 * h / c
    xs, vs = [], []
    dphi = radii[1] / radii[0] - 1.0  # :126
    for i in range(len(radii) - 1):  # :127-131
        r = radii[i]
        vphi = math.sqrt(GM) / math.sqrt(r)  # vphi_r :35-37
        phi = 0.0
        while phi < 2 * math.pi + 0.0:  # generate_circle! :109-119
            cx, sy = math.cos(phi), math.sin(phi)
            xs.append((0.0 + r * cx, 0.0 + r * sy, 0.0))
            vs.append((-vphi * sy, vphi * cx, 0.0))
            phi += dphi
        dphi = (radii[i + 1] - r) / r
    x = np.array(xs)
    v = np.array(vs)
    domain = geo.Rectangle(-bw, -bw, bw, bw)
    fields = {"v": 3, "a": 3, "P": 1, "rho": 1, "rho0": 1, "type": 1, "U": 1}
    init = {"x": x, "v": v, "type": np.zeros(len(x))}
    o_rho = ops.density_sum_fluid("wendland2", m, h, out="rho")
    o_rho0 = ops.density_sum_fluid("wendland2", m, h, out="rho0")
    o_p = ops.pressure_from_rho(c)
    o_f = ops.internal_force_lj("wendland2", m, h, dr_wall, E_wall, eps, rho0=rho0)
    o_ra = ops.fill("a", 0.0)
    o_rr = ops.fill("rho", 0.0)
    o_mv = ops.move_rev(dt)
    o_ac = ops.accelerate_rev_central(0.5 * dt, GM)

    def prologue(sys):  # :253-258
        sys.create_cell_list()
        sys.apply(o_rho0, self_=True)
        sys.apply(o_rho, self_=True)
        sys.apply(o_p)
        sys.apply(o_f)

    def step(sys):  # verlet_step! :220-230
        sys.apply(o_ac)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_rr)
        sys.apply(o_rho, self_=True)
        sys.apply(o_p)
        sys.apply(o_ra)
        sys.apply(o_f)
        sys.apply(o_ac)

    return Case("kepler_vortex", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, rho0=rho0, m=m, c=c, dt=dt, GM=GM, r0=r0, dr_wall=dr_wall, E_wall=E_wall,
                            eps=eps), dim=2)


# --------------------------------------------------------------------------- cylinder.jl
def cylinder(init) -> Case:
    """examples/cylinder.jl:29-61 (constants), :86-89 (make_system: the particles are imported from
    init/cylinder.vtp), :167-186 (the modified Verlet loop with the inflow buffer).  ``init`` is the path of that
    .vtp file or a mapping with ``x`` (n, 3) and ``type`` (n,) arrays (tests/golden/cylinder_init.npz holds the
    file's particles).  The step index k (t = k*dt, :178) is kept on the system object as ``step_index``."""
    chan_l, chan_w, cyl1, cyl_r = 2.2, 0.41, 0.2, 0.05
    dr = math.pi * cyl_r / 20
    h = 2.4 * dr
    bc_width = 6 * dr
    x2_min = -chan_w / 2 - 6 * dr
    x2_max = chan_w / 2 + 6 * dr
    U_max = 0.3
    rho0 = 1.0
    m0 = rho0 * dr ** 2
    c = 20.0 * U_max
    mu = 1.0e-3
    nu = 0.1 * h * c
    dt = 0.1 * h / c
    t_acc = 1.0
    FLUID, INFLOW, WALL, OBSTACLE = 0.0, 1.0, 2.0, 3.0
    if isinstance(init, str):
        from . import io as sp_io
        pts, pf = sp_io.read_vtp(init)
        x, typ = pts, pf["type"]
    else:
        x, typ = np.asarray(init["x"], dtype=np.float64), np.asarray(init["type"], dtype=np.float64)
    n = len(x)
    domain = geo.Rectangle(-bc_width, x2_min, chan_l, x2_max)
    fields = {"v": 3, "a": 3, "rho": 1, "Drho": 1, "P": 1, "m": 1, "type": 1}
    # Particle(x) :74-76: rho = rho0, m = m0; import_particles! then copies `type` from the file
    start = {"x": x, "rho": np.full(n, rho0), "m": np.full(n, m0), "type": typ}
    o_bom = ops.cyl_balance_of_mass("wendland2", h, nu)
    o_fp = ops.cyl_find_pressure(dt, c, rho0, -bc_width + h)
    o_if = ops.cyl_internal_force("wendland2", h, mu)
    o_mv = ops.move_types(dt, FLUID, INFLOW)
    o_ac = ops.cyl_accelerate(0.5 * dt, cyl1, U_max)

    def step(sys):  # :177-186
        k = getattr(sys, "step_index", 0) + 1
        sys.step_index = k
        t = k * dt
        sys.apply(o_ac)
        sys.apply(o_mv)
        sys.respawn("type", INFLOW, FLUID, 0.0, bc_width, rho=rho0, m=m0)  # add_new_particles! :145-156
        sys.apply(ops.set_inflow_speed(t, t_acc, U_max, chan_w, INFLOW))
        sys.create_cell_list()
        sys.apply(o_bom)
        sys.apply(o_fp)
        sys.apply(o_if)
        sys.apply(o_ac)

    return Case("cylinder", fields, domain, h, start, step,
                consts=dict(dr=dr, h=h, rho0=rho0, m0=m0, c=c, mu=mu, nu=nu, dt=dt, U_max=U_max, bc_width=bc_width,
                            chan_w=chan_w, t_acc=t_acc, OBSTACLE=OBSTACLE, L_char=0.1), dim=2)


def cylinder_force_coefficients(sys, consts) -> np.ndarray:
    """calculate_force, cylinder.jl:158-164: C = 2 F/(L_char U_mean^2), F = sum of m*a over the obstacle particles."""
    F = sys.reduce(K["SP_RED_FORCE_ON_TYPE"], ("a", "m", "type"), (consts["OBSTACLE"],), nout=3)
    U_mean = 2 / 3 * consts["U_max"]
    return 2.0 * np.asarray(F) / (consts["L_char"] * U_mean ** 2)


# --------------------------------------------------------------------------- rod.jl
def rod(dr: float = None) -> Case:
    """examples/rod.jl:17-41 (constants), :101-120 (make_geometry, force_computation!), :207-229 (Verlet loop): a
    clamped elastic rod pulled at its free end for pull_time, then left to vibrate.  A, H, B are RealMatrix fields
    (9 components).  The step index k (t = k*dt, :208) is kept on the system object as ``step_index``."""
    L, W, r_free = 5.0, 0.5, 1.0
    pull_force, pull_time = 1.0, 0.5
    c_l, c_s = 20.0, 200.0
    c_0 = math.sqrt(c_l ** 2 + 4 / 3 * c_s ** 2)
    rho0 = 1.0
    nu = 1.0e-4
    if dr is None:
        dr = W / 16
    h = 2.5 * dr
    vol = dr ** 2
    m = rho0 * vol
    dt = 0.1 * h / c_0
    grid = geo.Hexagrid(dr)
    body = geo.Rectangle(0.0, 0.0, L, W)
    domain = geo.Rectangle(-r_free, -r_free, L + r_free, W + r_free)
    x = geo.covering(grid, body)
    fields = {"v": 3, "f": 3, "X": 3, "A": 9, "H": 9, "B": 9, "e": 1}
    init = {"x": x, "X": x.copy()}
    o_A = ops.rod_find_A("wendland2", h)
    o_B = ops.rod_find_B(m, c_l, c_s)
    o_f = ops.rod_find_f("wendland2", h, m, vol, nu)
    o_pull = ops.rod_pull(L - h, (vol * pull_force) / (h * W))
    o_v = ops.rod_update_v(0.5 * dt, m, h)
    o_x = ops.rod_update_x(dt)

    def force_computation(sys, t):  # :112-120
        sys.apply(o_A)
        sys.apply(o_B)
        sys.apply(o_f)
        if t < pull_time:
            sys.apply(o_pull)

    def prologue(sys):  # :107-108
        sys.create_cell_list()
        force_computation(sys, 0.0)

    def step(sys):  # :207-208, :222-227
        k = getattr(sys, "step_index", 0)
        sys.step_index = k + 1
        t = k * dt
        sys.apply(o_v)
        sys.apply(o_x)
        sys.create_cell_list()
        force_computation(sys, t)
        sys.apply(o_v)

    return Case("rod", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, m=m, vol=vol, dt=dt, c_l=c_l, c_s=c_s, nu=nu, L=L, W=W, pull_time=pull_time),
                dim=2)


def rod_energy(sys, consts) -> float:
    """sum(p -> particle_energy(p), sys.particles), rod.jl:190-199, :213."""
    return float(sys.reduce(K["SP_RED_ENERGY_ROD"], ("v", "A"), (consts["m"], consts["c_s"], consts["c_l"]))[0])


# --------------------------------------------------------------------------- SHTC/ldc.jl
def shtc_ldc(N: int = 100, Re: float = 100.0) -> Case:
    """examples/SHTC/ldc.jl:16-39 (constants), :76-88 (geometry), :152-169 (loop): lid-driven cavity with the SHTC
    model — every particle carries a 3x3 distortion field A (relaxed by an RK4 step) and a stress tensor."""
    llid, vlid, rho0 = 1.0, 1.0, 1.0
    c_l, c_s = 20.0, 20.0
    tau = 6 * vlid * llid / (Re * c_s ** 2)
    dr = llid / N
    h = 2.4 * dr
    m = rho0 * dr ** 2
    acf = 1e-3
    wwall = 1.5 * h
    dt = 0.05 * h / c_l
    FLUID, WALL, LID = 0.0, 1.0, 2.0
    grid = geo.Hexagrid(dr)
    box = geo.Rectangle(0.0, 0.0, llid, llid)
    layer = geo.BoundaryLayer(box, grid, wwall)
    lid = geo.Specification(layer, geo.HalfSpace(1, ">", llid))
    walls = geo.Specification(layer, geo.HalfSpace(1, "<=", llid))
    domain = (walls + lid + box).boundarybox()
    xf, xl, xw = geo.covering(grid, box), geo.covering(grid, lid), geo.covering(grid, walls)
    x = np.concatenate([xf, xl, xw])
    n = len(x)
    typ = np.concatenate([np.full(len(xf), FLUID), np.full(len(xl), LID), np.full(len(xw), WALL)])
    v = np.zeros((n, 3))
    v[len(xf):len(xf) + len(xl), 0] = vlid
    A = np.tile(np.eye(3).ravel(), (n, 1))  # MAT1 (symmetric: row- and column-major agree)
    fields = {"v": 3, "rho": 1, "type": 1, "A": 9, "stress": 9}
    init = {"x": x, "v": v, "rho": np.full(n, rho0), "type": typ, "A": A}
    o_s = ops.shtc_find_stress(c_l, c_s, rho0, acf)
    o_v = ops.shtc_update_v("wendland2", h, dt, m)
    o_r = ops.shtc_update_rho("wendland2", h, dt, m)
    o_c = ops.shtc_convect_A("wendland2", h, dt, m, LID)
    o_x = ops.shtc_relax_A(dt, tau)
    o_m = ops.shtc_move(dt)

    def step(sys):  # :159-165
        sys.create_cell_list()
        sys.apply(o_s)
        sys.apply(o_v)
        sys.apply(o_r)
        sys.apply(o_c)
        sys.apply(o_x)
        sys.apply(o_m)

    return Case("shtc_ldc", fields, domain, h, init, step,
                consts=dict(dr=dr, h=h, m=m, dt=dt, tau=tau, c_l=c_l, c_s=c_s, rho0=rho0, acf=acf, vlid=vlid, LID=LID), dim=2)


# --------------------------------------------------------------------------- SHTC/beryllium.jl
def shtc_beryllium(dr: float = None) -> Case:
    """examples/SHTC/beryllium.jl:13-41 (constants, init_velocity), :109-127 (make_geometry with the J0/K0
    calibration), :230-245 (loop): a free beryllium plate vibrating in its first bending mode, SHTC solid model."""
    L, W = 0.06, 0.01
    c_s = 9046.59
    c_0 = c_s
    c_p = 4 * c_0
    c = math.sqrt(c_0 ** 2 + 4 / 3 * c_s ** 2)
    rho0 = 1845.0
    if dr is None:
        dr = W / 20
    h = 3.0 * dr
    m0 = rho0 * dr * dr
    dt = 0.05 * dr / c
    grid = geo.Hexagrid(dr)
    plate = geo.Rectangle(-L / 2, -W / 2, L / 2, W / 2)
    domain = geo.BoundaryLayer(plate, grid, W).boundarybox()
    x = geo.covering(grid, plate)
    n = len(x)
    # init_velocity :30-39
    Aamp, omega, alpha, a1, a2 = 4.3369e-5, 2.3597e5, 78.834, 56.6368, 57.6455
    sarg = alpha * (x[:, 0] + L / 2)
    v = np.zeros((n, 3))
    v[:, 1] = Aamp * omega * (a1 * (np.sinh(sarg) + np.sin(sarg)) - a2 * (np.cosh(sarg) + np.cos(sarg)))
    fields = {"m": 1, "v": 3, "P": 1, "f": 3, "A": 9, "T": 9, "L": 9, "J": 1, "K": 1, "J0": 1, "K0": 1}
    init = {"x": x, "m": np.full(n, m0), "v": v, "A": np.tile(np.eye(3).ravel(), (n, 1))}
    o_L = ops.be_find_L("wendland2", h, rho0)
    o_A = ops.be_update_A(0.5 * dt)
    o_J = ops.be_find_J("wendland2", h, rho0)
    o_T = ops.be_find_T(rho0, c_0, c_s)
    o_f = ops.be_find_f("wendland2", h, rho0, c_p)
    o_reset = ops.be_reset()
    o_v = ops.be_update_v(0.5 * dt)
    o_x = ops.advect(0.5 * dt)

    def prologue(sys):  # make_geometry :114-125
        sys.create_cell_list()
        sys.apply(o_J)
        sys.set("J0", 1.0 - sys.get("J"))   # the host loop `for p in sys.particles` of :117-120
        sys.set("K0", -sys.get("K"))
        sys.apply(o_reset)
        sys.apply(o_J)
        sys.apply(o_T)
        sys.apply(o_f)

    def step(sys):  # :232-244
        sys.apply(o_v)
        sys.apply(o_x)
        sys.create_cell_list()
        sys.apply(o_reset)
        sys.apply(o_L)
        sys.apply(o_A)
        sys.apply(o_x)
        sys.create_cell_list()
        sys.apply(o_reset)
        sys.apply(o_J)
        sys.apply(o_T)
        sys.apply(o_f)
        sys.apply(o_v)

    return Case("shtc_beryllium", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, m0=m0, dt=dt, rho0=rho0, c_0=c_0, c_s=c_s, c_p=c_p, L=L, W=W), dim=2)


def beryllium_energy(sys, consts) -> float:
    """E_kinetic + E_bulk + E_shear + E_penalty of beryllium.jl:189-204, evaluated on the host from downloaded fields
    (a per-frame diagnostic of the script)."""
    m, v, J, Kf = sys.get("m"), sys.get("v"), sys.get("J"), sys.get("K")
    A = sys.get("A").reshape(-1, 3, 3).transpose(0, 2, 1)
    G = A.transpose(0, 2, 1) @ A
    dev = G - (np.trace(G, axis1=1, axis2=2) / 3.0)[:, None, None] * np.eye(3)
    E = (0.5 * m * np.sum(v * v, axis=1) + 0.25 * m * consts["c_0"] ** 2 * ((1.0 - 1.0 / J) ** 2 + np.log(J) ** 2)
         + 0.25 * m * consts["c_s"] ** 2 * np.sum(dev * dev, axis=(1, 2)) + 0.5 * m * consts["c_p"] ** 2 * Kf ** 2)
    return float(np.sum(E))


# --------------------------------------------------------------------------- SHTC/twist3d.jl
def shtc_twist3d(dr: float = None) -> Case:
    """examples/SHTC/twist3d.jl:13-36 (constants, init_velocity), :102-120 (make_geometry), :242-254 (loop): a rubber
    column clamped below z = 0 and set spinning about its axis, SHTC solid in 3-D on a body-centred lattice."""
    H, W = 6.0, 1.0
    omega = 105.0
    rho0 = 1100.0
    Y, nu = 17e6, 0.495
    c_s = math.sqrt(0.5 / rho0 * Y / (1.0 + nu))
    c_0 = math.sqrt(nu * Y / (rho0 * (1.0 + nu) * (1.0 - 2 * nu)))
    c_p = c_0
    c = math.sqrt(c_0 ** 2 + 4 / 3 * c_s ** 2)
    if dr is None:
        dr = W / 24
    h = 3.0 * dr
    m0 = rho0 * dr * dr * dr
    dt = 0.2 * dr / c
    grid = geo.BodycenteredGrid(dr)
    column = geo.Box(-0.5 * W, -0.5 * W, -h, 0.5 * W, 0.5 * W, H + 0.1 * dr)
    domain = geo.Box(column.x1_min - W, column.x2_min - W, column.x3_min - W, column.x1_max + W, column.x2_max + W,
                     column.x3_max + W)   # boundarybox(column + BoundaryLayer(column, grid, W))
    x = geo.covering(grid, column)
    n = len(x)
    spin = (x[:, 2] > 0.0) * omega * np.sin(0.5 * math.pi * x[:, 2] / H)   # init_velocity :36-38
    v = np.column_stack([spin * x[:, 1], spin * -x[:, 0], spin * 0.0])
    fields = {"m": 1, "v": 3, "P": 1, "f": 3, "A": 9, "T": 9, "L": 9, "J": 1, "K": 1, "J0": 1, "K0": 1}
    init = {"x": x, "m": np.full(n, m0), "v": v, "A": np.tile(np.eye(3).ravel(), (n, 1))}
    o_L = ops.tw_find_L("wendland3", h, rho0)
    o_A = ops.tw_update_A(0.5 * dt)
    o_J = ops.tw_find_J("wendland3", h, rho0)
    o_T = ops.tw_find_T(rho0, c_0, c_s)
    o_f = ops.tw_find_f("wendland3", h, rho0, c_p)
    o_reset = ops.be_reset()
    o_v = ops.tw_update_v(0.5 * dt)
    o_x = ops.advect(0.5 * dt)

    def prologue(sys):  # :107-118
        sys.create_cell_list()
        sys.apply(o_J)
        sys.set("J0", 1.0 - sys.get("J"))
        sys.set("K0", -sys.get("K"))
        sys.apply(o_reset)
        sys.apply(o_J)
        sys.apply(o_T)
        sys.apply(o_f)

    def step(sys):  # :242-254
        sys.apply(o_v)
        sys.apply(o_x)
        sys.create_cell_list()
        sys.apply(o_reset)
        sys.apply(o_L)
        sys.apply(o_A)
        sys.apply(o_x)
        sys.create_cell_list()
        sys.apply(o_reset)
        sys.apply(o_J)
        sys.apply(o_T)
        sys.apply(o_f)
        sys.apply(o_v)

    return Case("shtc_twist3d", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, m0=m0, dt=dt, rho0=rho0, c_0=c_0, c_s=c_s, c_p=c_p, H=H, W=W, omega=omega), dim=3)


# --------------------------------------------------------------------------- SHTC/taco.jl
def shtc_taco(dr: float = None) -> Case:
    """examples/SHTC/taco.jl:11-37 (constants), :80-103 (make_geometry with the C_rho/C_lambda calibration),
    :242-255 (loop): Taylor-Couette flow between a resting inner and a rotating outer cylinder, SHTC fluid on a Vogel
    spiral.  The step index k (t = k*dt, :228) is kept on the system object as ``step_index``."""
    R1, R2, omega, Re = 1.0, 2.0, 1.0, 20.0
    c_s, c_0, rho0 = 30.0, 15.0, 1.0
    tau = 6 * omega * R2 * (R2 - R1) / (Re * c_s ** 2)
    if dr is None:
        dr = (R2 - R1) / 20
    h = 3.0 * dr
    wwall = 1.5 * h
    c_p = 0.01 * c_0
    c = math.sqrt(c_0 ** 2 + 4 / 3 * c_s ** 2)
    m0 = rho0 * dr * dr
    dt = 0.05 * dr / c
    FLUID, INNER, OUTER = 0.0, 1.0, 2.0
    grid = geo.VogelGrid(dr)
    fluid = geo.Circle(0.0, 0.0, R2) - geo.Circle(0.0, 0.0, R1)
    walls = geo.BoundaryLayer(fluid, grid, wwall)
    mid = 0.5 * (R1 + R2)
    inner = geo.Specification(walls, lambda X: np.sqrt(X[:, 0] * X[:, 0] + X[:, 1] * X[:, 1] + X[:, 2] * X[:, 2]) < mid)
    outer = geo.Specification(walls, lambda X: np.sqrt(X[:, 0] * X[:, 0] + X[:, 1] * X[:, 1] + X[:, 2] * X[:, 2]) > mid)
    domain = geo.BoundaryLayer(fluid, grid, 10 * wwall).boundarybox()
    xf, xi, xo = geo.covering(grid, fluid), geo.covering(grid, inner), geo.covering(grid, outer)
    x = np.concatenate([xf, xi, xo])
    n = len(x)
    typ = np.concatenate([np.full(len(xf), FLUID), np.full(len(xi), INNER), np.full(len(xo), OUTER)])
    fields = {"m": 1, "x0": 3, "v": 3, "P": 1, "f": 3, "A": 9, "T": 9, "L": 9, "rho": 1, "lambda": 1, "C_rho": 1,
              "C_lambda": 1, "type": 1}
    init = {"x": x, "x0": x.copy(), "m": np.full(n, m0), "A": np.tile(np.eye(3).ravel(), (n, 1)), "type": typ}
    o_L = ops.be_find_L("wendland2", h, 1.0)                       # find_L! :128-134 (ker = q.m*rDw)
    o_A = ops.be_update_A(0.5 * dt)                                # update_A! :136-139
    o_relax = ops.shtc_relax_A(dt, tau)                            # relax_A! :180-194
    o_rho = ops.be_find_J("wendland2", h, 1.0, J="rho", Kf="lambda")   # find_rho! :141-146, self = true
    o_T = ops.ta_find_T(rho0, c_0, c_s)
    o_f = ops.ta_find_f("wendland2", h, c_p, rho0)
    o_reset = ops.be_reset(J="rho", Kf="lambda", J0="C_rho", K0="C_lambda")   # reset! :164-170
    o_v = ops.ta_update_v(0.5 * dt, R1, R2, omega)

    def prologue(sys):  # :90-102
        sys.create_cell_list()
        sys.apply(o_rho, self_=True)
        sys.set("C_rho", rho0 - sys.get("rho"))
        sys.set("C_lambda", -sys.get("lambda"))
        sys.apply(o_reset)
        sys.apply(o_rho, self_=True)
        sys.apply(o_T)
        sys.apply(o_f)

    def step(sys):  # :227-228, :242-255
        k = getattr(sys, "step_index", 0)
        sys.step_index = k + 1
        t = k * dt
        sys.apply(o_v)
        sys.apply(ops.ta_update_x(0.5 * dt, omega, t + 0.5 * dt, OUTER))
        sys.create_cell_list()
        sys.apply(o_reset)
        sys.apply(o_L)
        sys.apply(o_A)
        sys.apply(o_relax)
        sys.apply(ops.ta_update_x(0.5 * dt, omega, t + dt, OUTER))
        sys.create_cell_list()
        sys.apply(o_reset)
        sys.apply(o_rho, self_=True)
        sys.apply(o_T)
        sys.apply(o_f)
        sys.apply(o_v)

    return Case("shtc_taco", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, m0=m0, dt=dt, rho0=rho0, c_0=c_0, c_s=c_s, c_p=c_p, tau=tau, R1=R1, R2=R2,
                            omega=omega, OUTER=OUTER), dim=2)


def taco_exact_velocity(x, consts):
    """vexact, taco.jl:39-42: the steady Couette profile."""
    R1, R2, omega = consts["R1"], consts["R2"], consts["omega"]
    r = np.sqrt(np.sum(x * x, axis=1))
    sc = R2 / r * (r / R1 - R1 / r) / (R2 / R1 - R1 / R2)
    return np.column_stack([sc * (-omega * x[:, 1]), sc * (omega * x[:, 0]), np.zeros(len(x))])


# --------------------------------------------------------------------------- collapse_dry_implicit.jl
def collapse_dry_implicit(dr: float = 1.0e-2) -> Case:
    """examples/collapse_dry_implicit.jl:47-114 (constants, make_system) and :218-233 (loop).
    The serial assemble_matrix + cg of :223-227 becomes the matrix-free CG of sp_poisson_cg."""
    dim = 2
    h = 2.8 * dr
    rho = 1000.0
    g = (0.0, -9.8, 0.0)
    mu = 8.4e-4
    m = dr ** dim * rho
    C_free = 10.0
    v_char = 5.0
    wcw, wch, bh, bw = 1.0, 2.0, 3.0, 4.0
    nlayers = 3.5
    dt = 0.1 * h / v_char
    grid = geo.Hexagrid(dr)
    box = geo.Rectangle(0.0, 0.0, bw, bh)
    fluid = geo.Rectangle(0.0, 0.0, wcw, wch)
    walls = geo.Specification(geo.BoundaryLayer(box, grid, 1.2 * dr), lambda X: X[:, 1] < bh)
    dummy = geo.Specification(geo.BoundaryLayer(box, grid, nlayers * dr) - walls, lambda X: X[:, 1] < bh)
    domain = (fluid + dummy + walls).boundarybox()
    xf, xw, xd = geo.covering(grid, fluid), geo.covering(grid, walls), geo.covering(grid, dummy)
    x = np.concatenate([xf, xw, xd])
    typ = np.concatenate([np.zeros(len(xf)), np.ones(len(xw)), np.full(len(xd), 2.0)])
    fields = {"v": 3, "Dv": 3, "P": 1, "div": 1, "L": 1, "lambda": 1, "type": 1, "b": 1}
    init = {"x": x, "type": typ}
    o_init = ops.isph_initialize(dt, g)
    o_visc = ops.isph_viscous_force("spline23", m, h, mu, rho)
    o_dll = ops.isph_div_L_lambda("spline23", m, h, rho, dim)
    o_b = ops.isph_projection_vector(h, dt)
    A = ops.isph_projection_matrix("spline23", m, h, rho, C_free)
    o_if = ops.isph_internal_force("spline23", m, h, rho)
    o_ac = ops.isph_accelerate(dt)

    def prologue(sys):  # :113
        sys.create_cell_list()

    def step(sys):  # :218-233
        sys.apply(o_init)
        sys.create_cell_list()
        sys.apply(o_visc)
        sys.apply(o_dll)
        sys.apply(o_b)
        sys.poisson_cg(A, "b", "P")
        sys.apply(o_if)
        sys.apply(o_ac)

    case = Case("collapse_dry_implicit", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, rho=rho, m=m, dt=dt, g=g, C_free=C_free, mu=mu), dim=2)
    case.ops = dict(init=o_init, visc=o_visc, dll=o_dll, b=o_b, A=A, force=o_if, acc=o_ac)
    return case


# --------------------------------------------------------------------------- test_collision_2d.jl
def collision_2d() -> Case:
    """tests/test_collision_2d.jl:14-59 (constants, make_system), :104-126 (verlet_step!, init)."""
    dr = 2.0e-2
    h = 2.4 * dr
    rho0 = 1000.0
    m = rho0 * dr ** 2
    c = 20.0
    circ_rad, dom_len, dom_wid, deltaX, deltaY = 0.4, 20.0, 20.0, 1.0, 0.2
    dt = 0.1 * h / c
    grid = geo.Squaregrid(dr)
    circ1 = geo.Circle(-0.5 * deltaX, -0.5 * deltaY, circ_rad)
    circ2 = geo.Circle(0.5 * deltaX, 0.5 * deltaY, circ_rad)
    domain = geo.Rectangle(-0.5 * dom_len, -0.5 * dom_wid, 0.5 * dom_len, 0.5 * dom_wid)
    x1, x2 = geo.covering(grid, circ1), geo.covering(grid, circ2)
    x = np.concatenate([x1, x2])
    v = np.zeros_like(x)
    v[: len(x1), 0] = 1.0
    v[len(x1):, 0] = -1.0
    fields = {"v": 3, "a": 3, "P": 1, "rho": 1, "rho0": 1}
    init = {"x": x, "v": v}
    o_rho = ops.density_sum("wendland2", m, h, out="rho")
    o_rho0 = ops.density_sum("wendland2", m, h, out="rho0")
    o_p = ops.pressure_from_rho(c)
    o_f = ops.internal_force_sym("wendland2", m, h, rho0)
    o_ra = ops.fill("a", 0.0)
    o_rr = ops.fill("rho", 0.0)
    o_mv = ops.advect(dt)
    o_ac = ops.kick(0.5 * dt)

    def prologue(sys):  # :121-126
        sys.create_cell_list()
        sys.apply(o_rho0, self_=True)
        sys.apply(o_rho, self_=True)
        sys.apply(o_p)
        sys.apply(o_f)

    def step(sys):  # verlet_step! :104-114
        sys.apply(o_ac)
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_rr)
        sys.apply(o_rho, self_=True)
        sys.apply(o_p)
        sys.apply(o_ra)
        sys.apply(o_f)
        sys.apply(o_ac)

    return Case("collision_2d", fields, domain, h, init, step, prologue,
                consts=dict(dr=dr, h=h, rho0=rho0, m=m, c=c, dt=dt, t_end=1.0), dim=2)


# --------------------------------------------------------------------------- synthetic S1 block
def lattice_box(n_side, jitter: float = 0.1, seed: int = 1234, shuffle: bool = True, dr: float = 1.0) -> Case:
    """n_side^3 (or (nx,ny,nz)) cubic lattice, h = 2 dr, positions jittered by
    U(-jitter*dr, jitter*dr) from Philox(seed), v ~ U(-0.01c, 0.01c) from Philox(seed+1), all fluid;
    domain = lattice bounds +- h; initial particle order shuffled with Philox(99) so the sort does work.
    The step is the collapse3d loop."""
    if isinstance(n_side, int):
        n_side = (n_side, n_side, n_side)
    nx, ny, nz = n_side
    h = 2.0 * dr
    rho0 = 1000.0
    m = rho0 * dr ** 3
    c = 50.0
    g = (0.0, 0.0, -9.8)
    mu = 8.4e-4
    nu = 1.0e-4
    dt = 0.1 * h / c
    I, J, Kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    x = np.empty((nx * ny * nz, 3))
    x[:, 0] = I.ravel() * dr
    x[:, 1] = J.ravel() * dr
    x[:, 2] = Kk.ravel() * dr
    n = len(x)
    if jitter:
        rng = np.random.Generator(np.random.Philox(seed))
        x += rng.uniform(-jitter * dr, jitter * dr, size=(n, 3))
    rng = np.random.Generator(np.random.Philox(seed + 1))
    v = rng.uniform(-0.01 * c, 0.01 * c, size=(n, 3))
    if shuffle:
        perm = np.random.Generator(np.random.Philox(99)).permutation(n)
        x, v = x[perm], v[perm]
    domain = geo.Box(-h, -h, -h, (nx - 1) * dr + h, (ny - 1) * dr + h, (nz - 1) * dr + h)
    fields = {"v": 3, "Dv": 3, "P": 1, "rho": 1, "Drho": 1, "type": 1}
    init = {"x": x, "v": v, "rho": np.full(n, rho0), "type": np.zeros(n)}
    o_bom = ops.balance_of_mass("wendland3", m, h, nu)
    o_fp = ops.find_pressure(dt, c, rho0)
    o_if = ops.internal_force("wendland3", m, h, mu, rho0)
    o_mv = ops.move(dt)
    o_ac = ops.accelerate(0.5 * dt, g)

    def step(sys):
        sys.apply(o_mv)
        sys.create_cell_list()
        sys.apply(o_bom)
        sys.apply(o_fp)
        sys.apply(o_if)
        sys.apply(o_ac)
        sys.apply(o_ac)

    case = Case(f"lattice_box_{nx}x{ny}x{nz}", fields, domain, h, init, step,
                consts=dict(dr=dr, h=h, rho0=rho0, m=m, c=c, mu=mu, nu=nu, dt=dt, g=g),
                program=K["SP_PROGRAM_WCSPH_3D"], program_fields=("x", "v", "Dv", "rho", "Drho", "P", "type"),
                program_params=(float(K["SP_KERNEL_WENDLAND3"]), m, h, 2 * nu, dt, c * c, rho0, mu, *g), dim=3)
    case.ops = dict(bom=o_bom, fp=o_fp, force=o_if, move=o_mv, acc=o_ac)
    return case
