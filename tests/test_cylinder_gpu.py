"""GPU parity for examples/cylinder.jl: inflow buffer (sp_respawn), per-particle-mass operators, the obstacle force
reduction and a stretch of the time loop with particles entering and leaving, through the C ABI against the oracle."""
import numpy as np
import pytest

import smoothedparticles_jl_b200 as sp
from smoothedparticles_jl_b200 import ParticleSystem, configs, geometry as geo, operators as ops
from oracle.oracle import OracleSystem
from parity import RTOL_STEP, assert_fields_close, neighbour_sets_equal
from test_cylinder_cpu import cylinder_init

pytestmark = pytest.mark.gpu
K = sp.K


def test_respawn_on_device_keeps_the_reference_order():
    rng = np.random.default_rng(11)
    n = 5000
    x = rng.uniform(-1, 1, (n, 3))
    typ = rng.integers(0, 3, n).astype(np.float64)
    x[::9, 0] = 0.0
    v = rng.uniform(-1, 1, (n, 3))
    fields = {"v": 3, "rho": 1, "m": 1, "type": 1}
    box = geo.Box(-2.0, -2.0, -2.0, 2.0, 2.0, 2.0)
    dev, ora = ParticleSystem(fields, box, 0.3), OracleSystem(fields, box, 0.3)
    for s in (dev, ora):
        s.add_particles(x=x, v=v, rho=np.full(n, 3.0), m=np.full(n, 0.5), type=typ)
        s.create_cell_list()  # the device slots are now sorted by cell: the new particles must still follow the
        #                       reference (index) order of their sources
    a, b = dev.respawn("type", 1.0, 0.0, 0.0, 0.25, rho=7.0, m=0.125), ora.respawn("type", 1.0, 0.0, 0.0, 0.25, rho=7.0, m=0.125)
    assert a == b > 0 and len(dev) == len(ora) == n + a
    for f in ("x", "v", "rho", "m", "type"):
        assert np.array_equal(dev.get(f), ora.get(f)), f
    # a second round right away (no cell list in between), then a rebuild
    for s in (dev, ora):
        s.set("type", np.where(s.get("x")[:, 0] < -0.5, 1.0, s.get("type")))
    a, b = dev.respawn("type", 1.0, 2.0, -0.9, 0.5), ora.respawn("type", 1.0, 2.0, -0.9, 0.5)
    assert a == b > 0
    for s in (dev, ora):
        s.create_cell_list()
    assert len(dev) == len(ora)
    for f in ("x", "v", "rho", "m", "type"):
        assert np.array_equal(dev.get(f), ora.get(f)), f
    assert np.array_equal(dev.cell_keys(), ora.cell_keys())
    assert dev.respawn("type", 5.0, 0.0, 0.0, 0.25) == 0


def test_cylinder_operators_single_call():
    case = configs.cylinder(cylinder_init())
    c = case.consts
    rng = np.random.default_rng(2)
    n = case.n
    v = rng.uniform(-0.3, 0.3, (n, 3)) * np.array([1, 1, 0])
    rho = rng.uniform(0.95, 1.05, n)
    P = rng.uniform(-1, 1, n)
    m = c["m0"] * rng.uniform(0.8, 1.2, n)
    dev, ora = case.make(ParticleSystem), case.make(OracleSystem)
    for s in (dev, ora):
        s.set("v", v)
        s.set("rho", rho)
        s.set("P", P)
        s.set("m", m)
        s.create_cell_list()
    assert len(dev) == len(ora)
    assert np.array_equal(dev.cell_keys(), ora.cell_keys())
    assert neighbour_sets_equal(dev, ora, ordered=True)
    for strict, rtol in ((False, RTOL_STEP), (True, 1e-12)):
        for s in (dev, ora):
            s.apply(ops.fill("Drho", 0.0))
            s.apply(ops.fill("a", 0.0))
            s.apply(ops.cyl_balance_of_mass("wendland2", c["h"], c["nu"]), strict_order=strict)
            s.apply(ops.cyl_internal_force("wendland2", c["h"], c["mu"]), strict_order=strict)
        assert_fields_close(dev, ora, ["Drho", "a"], rtol=rtol, what=f"cylinder pair operators strict={strict}")
    F = dev.reduce(K["SP_RED_FORCE_ON_TYPE"], ("a", "m", "type"), (3.0,), nout=3)
    Fo = ora.reduce(K["SP_RED_FORCE_ON_TYPE"], ("a", "m", "type"), (3.0,), nout=3)
    assert np.max(np.abs(F - Fo)) <= 1e-12 * np.max(np.abs(Fo)) and np.max(np.abs(Fo)) > 0
    for s in (dev, ora):
        s.apply(ops.cyl_accelerate(0.5 * c["dt"], 0.2, c["U_max"]))
        s.apply(ops.set_inflow_speed(0.4, c["t_acc"], c["U_max"], c["chan_w"]))
        s.apply(ops.move_types(c["dt"], 0.0, 1.0))
        s.apply(ops.cyl_find_pressure(c["dt"], c["c"], c["rho0"], -c["bc_width"] + c["h"]))
    assert_fields_close(dev, ora, ["x", "v", "a", "rho", "Drho", "P"], rtol=1e-14, what="cylinder unary operators")
    vd = dev.get("v")
    assert np.array_equal(vd[dev.get("type") == 1.0], ora.get("v")[ora.get("type") == 1.0])   # inflow profile: exact


def test_cylinder_time_loop_with_inflow_and_outflow():
    case = configs.cylinder(cylinder_init())
    c = case.consts
    dev, ora = case.make(ParticleSystem), case.make(OracleSystem)
    n0 = len(dev)
    dev.step_index = ora.step_index = 3300   # full inflow speed (t > t_acc)
    for k in range(150):
        case.step(dev)
        case.step(ora)
        assert len(dev) == len(ora), f"particle counts differ at step {k}"
    assert len(dev) != n0 and dev.n_removed == ora.n_removed > 0
    assert np.array_equal(dev.get("type"), ora.get("type"))
    assert_fields_close(dev, ora, ["x", "v", "rho", "P", "a"], rtol=1e-9, what="cylinder 150 steps",
                        floors={"P": c["c"] ** 2 * c["rho0"] * 1e-3})
    Cd, Co = configs.cylinder_force_coefficients(dev, c), configs.cylinder_force_coefficients(ora, c)
    assert np.max(np.abs(Cd - Co)) <= 1e-8 * max(np.max(np.abs(Co)), 1e-3)
    ora.set("x", dev.get("x"))
    dev.create_cell_list()
    ora.create_cell_list()
    assert len(dev) == len(ora)
    assert np.array_equal(dev.cell_keys(), ora.cell_keys())
    assert neighbour_sets_equal(dev, ora, ordered=True)
