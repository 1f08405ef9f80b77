"""Oracle pins for examples/cylinder.jl: the initial-state fixture against the reference's own .vtp file summary,
the inflow buffer (add_new_particles!, :145-156) against a literal Python loop, the per-particle-mass operators
against an O(N^2) numpy evaluation, and a stretch of the script's time loop."""
import hashlib
import json
import os

import numpy as np

import smoothedparticles_jl_b200 as sp
from smoothedparticles_jl_b200 import configs, geometry as geo, operators as ops
from oracle.oracle import OracleSystem

K = sp.K
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cylinder_init():
    d = np.load(os.path.join(GOLDEN, "cylinder_init.npz"))
    xy = d["xy"]
    return {"x": np.column_stack([xy, np.zeros(len(xy))]), "type": d["type"].astype(np.float64)}


def rDwendland2(h, r):  # kernels.jl:140-147
    x = r / h
    return np.where(x <= 1.0, -44.563384065730695 * (1 - x) ** 3 / h ** 4, 0.0)


def test_cylinder_fixture_is_the_reference_file():
    # tests/golden/cylinder_vtp_summary.json was computed from examples/init/cylinder.vtp (tools/make_vtp_golden.py)
    summary = json.load(open(os.path.join(GOLDEN, "cylinder_vtp_summary.json")))
    init = cylinder_init()
    assert len(init["x"]) == summary["n"] == 16912
    assert hashlib.sha256(np.ascontiguousarray(init["x"]).tobytes()).hexdigest() == summary["points_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(init["type"]).tobytes()).hexdigest() == summary["fields"]["type"]["sha256"]


def respawn_literal(x, typ, fields, from_type, to_type, x1_min, shift, constants):
    """cylinder.jl:145-156, line by line."""
    new_x, new_f = [], {k: [] for k in fields}
    typ = typ.copy()
    for i in range(len(x)):
        if typ[i] == from_type and x[i, 0] >= x1_min:
            typ[i] = to_type
            new_x.append(x[i] - shift * np.array([1.0, 0.0, 0.0]))
            for k in fields:
                new_f[k].append(constants.get(k, 0.0))
    return typ, np.array(new_x).reshape(-1, 3), new_f


def test_respawn_matches_the_literal_loop():
    rng = np.random.default_rng(11)
    n = 500
    x = rng.uniform(-1, 1, (n, 3))
    typ = rng.integers(0, 3, n).astype(np.float64)
    x[::9, 0] = 0.0  # the threshold itself is inside (>=)
    s = OracleSystem({"v": 3, "rho": 1, "m": 1, "type": 1}, geo.Box(-2.0, -2.0, -2.0, 2.0, 2.0, 2.0), 0.3)
    s.add_particles(x=x, v=rng.uniform(-1, 1, (n, 3)), rho=np.full(n, 3.0), m=np.full(n, 0.5), type=typ)
    want_t, want_x, _ = respawn_literal(x, typ, [], 1.0, 0.0, 0.0, 0.25, {})
    added = s.respawn("type", 1.0, 0.0, 0.0, 0.25, rho=7.0, m=0.125)
    assert added == len(want_x) > 0
    assert len(s) == n + added
    assert np.array_equal(s.get("type")[:n], want_t)
    assert np.array_equal(s.get("x")[n:], want_x)
    assert np.all(s.get("type")[n:] == 1.0) and np.all(s.get("rho")[n:] == 7.0) and np.all(s.get("m")[n:] == 0.125)
    assert np.all(s.get("v")[n:] == 0.0)
    assert s.respawn("type", 5.0, 0.0, 0.0, 0.25) == 0


def test_cylinder_pair_operators_against_brute_force():
    case = configs.cylinder(cylinder_init())
    c = case.consts
    s = case.make(OracleSystem)
    keep = np.flatnonzero(case.init["x"][:, 0] < 0.45)  # inflow buffer, walls, obstacle and the fluid around it
    s = OracleSystem(case.fields, case.domain, case.h)
    rng = np.random.default_rng(2)
    n = len(keep)
    x, typ = case.init["x"][keep], case.init["type"][keep]
    v = rng.uniform(-0.3, 0.3, (n, 3)) * np.array([1, 1, 0])
    rho = rng.uniform(0.95, 1.05, n)
    P = rng.uniform(-1, 1, n)
    m = c["m0"] * rng.uniform(0.8, 1.2, n)
    s.add_particles(x=x, v=v, rho=rho, P=P, m=m, type=typ)
    s.create_cell_list()
    assert len(s) == n
    s.apply(ops.cyl_balance_of_mass("wendland2", c["h"], c["nu"]))
    s.apply(ops.cyl_internal_force("wendland2", c["h"], c["mu"]))
    d = x[:, None, :] - x[None, :, :]
    r = np.sqrt(np.sum(d * d, axis=2))
    nb = (r <= c["h"]) & ~np.eye(n, dtype=bool)
    dv = v[:, None, :] - v[None, :, :]
    ker = np.where(nb, m[None, :] * rDwendland2(c["h"], r), 0.0)
    xv = np.sum(d * dv, axis=2)
    ff = nb & (typ[:, None] == 0.0) & (typ[None, :] == 0.0)
    want_D = np.sum(ker * xv, axis=1) + np.sum(np.where(ff, 2 * c["nu"] / rho[:, None] * (rho[:, None] - rho[None, :]), 0.0),
                                               axis=1)
    assert np.max(np.abs(s.get("Drho") - want_D)) <= 1e-12 * np.max(np.abs(want_D))
    pr = P / rho ** 2
    cp = -ker * (pr[:, None] + pr[None, :])
    cv = 8.0 * ker * c["mu"] / (rho[:, None] * rho[None, :]) * xv / (r * r + 0.01 * c["h"] * c["h"])
    want_a = np.sum((cp + cv)[:, :, None] * d, axis=1)
    assert np.max(np.abs(s.get("a") - want_a)) <= 1e-12 * np.max(np.abs(want_a))
    # unary operators
    x0, v0, a0 = s.get("x"), s.get("v"), s.get("a")
    s.apply(ops.cyl_accelerate(0.5 * c["dt"], 0.2, c["U_max"]))
    f = np.column_stack([0.2 - x0[:, 0], -x0[:, 1], np.zeros(n)])
    absf2 = (0.2 - x0[:, 0]) ** 2 + x0[:, 1] ** 2
    want_v = np.where((typ == 0.0)[:, None], v0 + 0.5 * c["dt"] * (a0 + 0.3 * c["U_max"] ** 2 * f / absf2[:, None]), v0)
    np.testing.assert_allclose(s.get("v"), want_v, rtol=1e-15, atol=1e-18)
    s.apply(ops.set_inflow_speed(0.4, c["t_acc"], c["U_max"], c["chan_w"]))
    v1 = 0.4 * c["U_max"] * (1.0 - (2.0 * x0[:, 1] / c["chan_w"]) ** 2)
    got = s.get("v")
    assert np.array_equal(got[typ == 1.0, 0], v1[typ == 1.0]) and np.all(got[typ == 1.0, 1:] == 0.0)
    assert np.array_equal(got[typ != 1.0], want_v[typ != 1.0]) or np.allclose(got[typ != 1.0], want_v[typ != 1.0], rtol=1e-15)
    s.apply(ops.move_types(c["dt"], 0.0, 1.0))
    moves = (typ == 0.0) | (typ == 1.0)
    np.testing.assert_allclose(s.get("x"), np.where(moves[:, None], x0 + c["dt"] * got, x0), rtol=1e-15, atol=0)
    assert np.all(s.get("a") == 0.0)
    rho_before, D = s.get("rho"), s.get("Drho")
    s.apply(ops.cyl_find_pressure(c["dt"], c["c"], c["rho0"], -c["bc_width"] + c["h"]))
    xx = s.get("x")[:, 0]
    want_rho = np.where(xx >= -c["bc_width"] + c["h"], rho_before + D * c["dt"], rho_before)
    assert np.array_equal(s.get("rho"), want_rho)
    assert np.array_equal(s.get("P"), c["c"] ** 2 * (want_rho - c["rho0"])) and np.all(s.get("Drho") == 0.0)


def test_cylinder_time_loop_with_inflow():
    case = configs.cylinder(cylinder_init())
    c = case.consts
    s = case.make(OracleSystem)
    s.step_index = 3300  # t > t_acc: the inflow runs at full speed (the script reaches this after 3 300 steps)
    n0 = len(s)
    added = []
    respawn = s.respawn
    s.respawn = lambda *a, **k: added.append(respawn(*a, **k)) or added[-1]
    for _ in range(150):
        case.step(s)
    removed = s.n_removed  # cumulative; the outflow: fluid leaving the domain at x > chan_l is dropped by create_cell_list!
    assert sum(added) > 0, "no particle left the inflow buffer: the respawn path was not exercised"
    assert len(s) == n0 + sum(added) - removed
    typ = s.get("type")
    # every particle that left the buffer was replaced: the number of INFLOW particles is constant
    assert np.sum(typ == 1.0) == np.sum(case.init["type"] == 1.0)
    assert np.sum(typ == 0.0) == np.sum(case.init["type"] == 0.0) + sum(added) - removed
    assert np.sum(typ >= 2.0) == np.sum(case.init["type"] >= 2.0)
    x, v = s.get("x"), s.get("v")
    assert np.all(x[typ == 1.0, 0] < 0.0) and np.all(x[typ == 1.0, 0] >= -c["bc_width"] - 1e-12)
    assert np.all(np.isfinite(v)) and np.max(np.abs(v)) < c["c"]
    C = configs.cylinder_force_coefficients(s, c)
    assert np.all(np.isfinite(C)) and C.shape == (3,) and C[2] == 0.0
    # walls and obstacle never move (removal may renumber them: compare as sets of coordinates)
    fixed0 = case.init["x"][case.init["type"] >= 2.0]
    fixed1 = x[typ >= 2.0]
    assert np.array_equal(fixed0[np.lexsort(fixed0.T)], fixed1[np.lexsort(fixed1.T)])


def test_max_speed_reduction_on_the_oracle():
    case = configs.cylinder(cylinder_init())
    s = case.make(OracleSystem)
    rng = np.random.default_rng(1)
    v = rng.uniform(-2, 2, (case.n, 3))
    s.set("v", v)
    want = np.max(np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]))
    assert s.reduce(K["SP_RED_MAX_SPEED"], ("v",), (), nout=1)[0] == want
    assert sp.cfl_time_step(s, 0.1, case.h, 6.0) == 0.1 * case.h / (6.0 + want)
