"""GPU parity tests for the reversible (fixed-point) WCSPH operators of examples/collapse_symplectic.jl,
examples/Kepler_vortex.jl and examples/utils/FixPA.jl, through the C ABI against the CPU oracle, plus the
size-independent property the scripts exist to demonstrate: reverting the velocities retraces the run exactly.
"""
import numpy as np
import pytest

import smoothedparticles_jl_b200 as sp
from smoothedparticles_jl_b200 import ParticleSystem, configs, geometry as geo, operators as ops
from oracle.oracle import OracleSystem
from parity import RTOL_STEP, assert_fields_close, neighbour_sets_equal
from test_symplectic_cpu import TWO30, rev_add_exact, symplectic_energy

pytestmark = pytest.mark.gpu
K = sp.K


def _pair(case):
    return case.make(ParticleSystem), case.make(OracleSystem)


def test_rev_add_on_device_is_exact_integer_arithmetic():
    # FixPA.jl:11-42 — bit-exact against Python integers, incl. rounding ties and lattice values
    rng = np.random.default_rng(5)
    n = 4000
    x = rng.uniform(-3, 3, (n, 3))
    v = rng.uniform(-50, 50, (n, 3))
    x[:50, 0] = (rng.integers(-10 ** 9, 10 ** 9, 50) + 0.5) / TWO30
    x[50:100, 1] = rng.integers(-10 ** 9, 10 ** 9, 50) / TWO30
    a = rng.uniform(-20, 20, (n, 3))
    typ = np.zeros(n)
    typ[::7] = 1.0
    dt = 6e-5
    s = ParticleSystem({"v": 3, "a": 3, "type": 1}, geo.Box(-10.0, -10.0, -10.0, 10.0, 10.0, 10.0), 1.0)
    s.add_particles(x=x, v=v, a=a, type=typ)
    s.apply(ops.move_rev(dt))
    fl = typ[:, None] == 0.0
    assert np.array_equal(s.get("x"), np.where(fl, rev_add_exact(x, dt * v), x))
    g = (0.0, -9.8, 0.0)
    s.apply(ops.accelerate_rev(0.5 * dt, g))
    assert np.array_equal(s.get("v"), np.where(fl, rev_add_exact(v, (0.5 * dt) * (a + np.asarray(g))), v))
    GM = 1000.0
    x1, v1 = s.get("x"), s.get("v")
    s.apply(ops.accelerate_rev_central(0.5 * dt, GM))
    nrm = np.sqrt((x1[:, 0] * x1[:, 0] + x1[:, 1] * x1[:, 1]) + x1[:, 2] * x1[:, 2])
    k = -GM / (nrm * nrm * nrm)
    want = np.where(fl, rev_add_exact(v1, (0.5 * dt) * rev_add_exact(a, k[:, None] * x1)), v1)
    assert np.array_equal(s.get("v"), want)


def test_symplectic_pair_operators_single_call():
    case = configs.collapse_symplectic(dr=2e-2)
    c = case.consts
    rng = np.random.default_rng(3)
    x = case.init["x"].copy()
    typ = case.init["type"]
    nf = int(np.sum(typ == 0.0))
    x[typ == 0.0] += rng.uniform(-0.3, 0.3, (nf, 3)) * c["dr"] * np.array([1, 1, 0])  # into the LJ range of the walls
    P = rng.uniform(0, 1e4, len(x))
    rho = np.where(typ == 0.0, rng.uniform(900, 1100, len(x)), 0.0)
    dev, ora = _pair(case)
    for s in (dev, ora):
        s.set("x", x)
        s.set("P", P)
        s.set("rho", rho)
        s.create_cell_list()
    assert np.array_equal(dev.cell_keys(), ora.cell_keys())
    assert neighbour_sets_equal(dev, ora, ordered=True)
    o_rho0 = ops.density_sum_fluid("wendland2", c["m"], c["h"], out="rho0")
    o_f = ops.internal_force_lj("wendland2", c["m"], c["h"], c["dr_wall"], c["E_wall"], c["eps"])
    o_fk = ops.internal_force_lj("wendland2", c["m"], c["h"], c["dr_wall"], c["E_wall"], c["eps"], rho0=c["rho0"])
    o_U = ops.lj_potential(c["h"], c["m"], c["E_wall"], c["dr_wall"], c["eps"])
    for strict, rtol in ((False, RTOL_STEP), (True, 1e-12)):
        for s in (dev, ora):
            for f in ("rho0", "a", "U"):
                s.apply(ops.fill(f, 0.0))
            s.apply(o_rho0, self_=True, strict_order=strict)
            s.apply(o_f, strict_order=strict)
            s.apply(o_U, strict_order=strict)
        assert np.max(np.abs(ora.get("U"))) > 0, "the LJ branch was not exercised"
        assert_fields_close(dev, ora, ["rho0", "a", "U"], rtol=rtol, what=f"symplectic single calls strict={strict}")
        assert np.all(dev.get("a")[typ != 0.0] == 0.0) and np.all(dev.get("rho0")[typ != 0.0] == 0.0)
        for s in (dev, ora):
            s.apply(ops.fill("a", 0.0))
            s.apply(o_fk, strict_order=strict)
        assert_fields_close(dev, ora, ["a"], rtol=rtol, what=f"Kepler pressure form strict={strict}")


@pytest.mark.parametrize("maker,kw,nsteps", [(configs.collapse_symplectic, dict(dr=2e-2), 25),
                                             (configs.kepler_vortex, dict(N_rings=12), 25)])
def test_symplectic_time_loop_parity(maker, kw, nsteps):
    case = maker(**kw)
    c = case.consts
    dev, ora = _pair(case)
    case.prologue(dev)
    case.prologue(ora)
    assert_fields_close(dev, ora, ["rho0", "rho", "P", "a"], what=f"{case.name} prologue",
                        floors={"P": c["c"] ** 2 * c["rho0"] * 1e-6})
    for _ in range(nsteps):
        case.step(dev)
        case.step(ora)
    assert len(dev) == len(ora) == case.n
    # x and v live on the 2^-30 lattice: device and oracle agree bit for bit except where a 1e-16 difference in the
    # acceleration flips one rounding (one lattice step, 9.3e-10)
    xd, xo, vd, vo = dev.get("x"), ora.get("x"), dev.get("v"), ora.get("v")
    assert np.max(np.abs(xd - xo)) <= 4 / TWO30
    assert np.max(np.abs(vd - vo)) <= 64 / TWO30
    assert np.mean(xd != xo) < 1e-3 and np.mean(vd != vo) < 1e-2
    # a flipped rounding moves a particle by 2^-30 and its neighbours' density by ~1e-8 relative: the bar after N
    # steps is the amplification of that, not the per-call bar (which the prologue above and the single-call test hold)
    assert_fields_close(dev, ora, ["rho", "a"], rtol=1e-5, what=f"{case.name} {nsteps} steps",
                        floors={"a": min(1.0, c["c"] ** 2 / c["h"])})
    ora.set("x", xd)
    dev.create_cell_list()
    ora.create_cell_list()
    assert np.array_equal(dev.cell_keys(), ora.cell_keys())
    assert neighbour_sets_equal(dev, ora, ordered=True)


@pytest.mark.parametrize("maker,kw,nsteps", [(configs.collapse_symplectic, dict(dr=2e-2), 60),
                                             (configs.collapse_symplectic, dict(dr=5e-3), 40),
                                             (configs.kepler_vortex, dict(), 30)])
def test_reverting_velocities_retraces_the_run_exactly_on_device(maker, kw, nsteps):
    # collapse_symplectic.jl:232-255 (revert = true), see tests/test_symplectic_cpu.py: exact, at any size
    case = maker(**kw)
    s = case.make(ParticleSystem)
    case.prologue(s)
    case.step(s)
    x1, v1, a1 = s.get("x"), s.get("v"), s.get("a")
    fl = s.get("type") == 0.0
    assert np.array_equal(x1[fl] * TWO30, np.rint(x1[fl] * TWO30))
    for _ in range(nsteps):
        case.step(s)
    assert len(s) == case.n
    assert np.max(np.abs(s.get("x") - x1)) > 100 / TWO30, "nothing moved: the test would be vacuous"
    s.set("v", -s.get("v"))
    for _ in range(nsteps):
        case.step(s)
    assert np.array_equal(s.get("x"), x1)
    assert np.array_equal(s.get("v"), -v1)
    assert np.array_equal(s.get("a"), a1)


def test_collapse_symplectic_energy_on_device():
    case = configs.collapse_symplectic(dr=2e-2)
    c = case.consts
    dev, ora = _pair(case)
    o_U = ops.lj_potential(c["h"], c["m"], c["E_wall"], c["dr_wall"], c["eps"])

    def energy(s):
        s.apply(ops.fill("U", 0.0))
        s.apply(o_U)
        return symplectic_energy(s, c)

    case.prologue(dev)
    case.prologue(ora)
    E0 = energy(dev)
    assert abs(E0 - energy(ora)) <= 1e-10 * abs(E0)
    for _ in range(200):
        case.step(dev)
    v = dev.get("v")
    kin = 0.5 * c["m"] * float(np.sum(v * v))
    drift = abs(energy(dev) - E0)
    # oracle at the same resolution: drift = 7 % of the kinetic energy gained = 1.7e-4 of |E0| after 200 steps
    assert kin > 0 and drift < 0.2 * kin and drift < 1e-3 * abs(E0)


def test_zoo_operators_reject_the_experimental_kernels_loudly():
    # operators beyond the BASELINE configs are built for the cached-list and strict-order kernels only
    case = configs.collapse_symplectic(dr=4e-2)
    c = case.consts
    s = case.make(ParticleSystem)
    s.create_cell_list()
    op = ops.density_sum_fluid("wendland2", c["m"], c["h"], out="rho")
    s.apply(op, self_=True)
    s.apply(op, self_=True, strict_order=True)
    with pytest.raises(sp.SpError):
        s.apply(op, self_=True, tile_kernel=True)
