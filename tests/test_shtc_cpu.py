"""Oracle pins for examples/SHTC/ldc.jl (SHTC fluid with a full 3x3 distortion field): every operator against an
independent numpy evaluation of the script's formulas; the order-dependent convect_A! against a literal sequential
loop over the oracle's own visiting order; a stretch of the time loop."""
import numpy as np

import smoothedparticles_jl_b200 as sp
from smoothedparticles_jl_b200 import configs, operators as ops
from oracle.oracle import OracleSystem

K = sp.K


def rDwendland2(h, r):
    x = r / h
    return np.where(x <= 1.0, -44.563384065730695 * (1 - x) ** 3 / h ** 4, 0.0)


def mat(a):
    """(n, 9) column-major RealMatrix field -> (n, 3, 3)."""
    return a.reshape(-1, 3, 3).transpose(0, 2, 1)


def dev(G):
    return G - np.trace(G, axis1=-2, axis2=-1)[..., None, None] / 3.0 * np.eye(3)


def corner_patch(seed=6):
    """The top-left corner of the cavity (fluid, wall and lid particles) with perturbed v, rho and A."""
    case = configs.shtc_ldc()
    keep = np.flatnonzero((case.init["x"][:, 0] < 0.22) & (case.init["x"][:, 1] > 0.8))
    rng = np.random.default_rng(seed)
    n = len(keep)
    x = case.init["x"][keep] + rng.uniform(-0.1, 0.1, (n, 3)) * case.consts["dr"] * np.array([1, 1, 0])
    typ = case.init["type"][keep]
    v = rng.uniform(-1, 1, (n, 3)) * np.array([1, 1, 0])
    rho = rng.uniform(0.97, 1.03, n)
    A = np.tile(np.eye(3), (n, 1, 1)) + rng.uniform(-0.05, 0.05, (n, 3, 3))
    A[:, 2, :2] = 0.0
    A[:, :2, 2] = 0.0
    s = OracleSystem(case.fields, case.domain, case.h)
    s.add_particles(x=x, v=v, rho=rho, type=typ, A=A.transpose(0, 2, 1).reshape(n, 9))
    s.create_cell_list()
    assert len(s) == n and set(np.unique(typ)) == {0.0, 1.0, 2.0}
    return case, s


def test_shtc_operators_against_numpy():
    case, s = corner_patch()
    c = case.consts
    h, dtm = c["h"], c["dt"] * c["m"]
    x, v, rho, typ, A = s.get("x"), s.get("v"), s.get("rho"), s.get("type"), mat(s.get("A"))
    n = len(x)
    d = x[:, None, :] - x[None, :, :]
    r = np.sqrt(np.sum(d * d, axis=2))
    nb = (r <= h) & ~np.eye(n, dtype=bool)
    rD = np.where(nb, rDwendland2(h, r), 0.0)
    # find_stress!  ldc.jl:118-121
    s.apply(ops.shtc_find_stress(c["c_l"], c["c_s"], c["rho0"], c["acf"]))
    G = A.transpose(0, 2, 1) @ A
    S = (c["c_l"] ** 2 * (rho - c["rho0"] / (1.0 + c["acf"])))[:, None, None] * np.eye(3) \
        + (c["c_s"] ** 2 * rho)[:, None, None] * (G @ dev(G))
    got = mat(s.get("stress"))
    assert np.max(np.abs(got - S)) <= 1e-12 * np.max(np.abs(S))
    # update_v!  ldc.jl:123-127
    s.apply(ops.shtc_update_v("wendland2", h, c["dt"], c["m"]))
    Sr = got / (rho * rho)[:, None, None]
    T = Sr[:, None] + Sr[None, :]                                      # (p, q, 3, 3)
    dv = np.einsum("pq,pqij,pqj->pi", -dtm * rD, T, d)
    want_v = np.where((typ == 0.0)[:, None], v + dv, v)
    assert np.max(np.abs(s.get("v") - want_v)) <= 1e-12 * np.max(np.abs(want_v))
    # update_rho!  ldc.jl:90-94 (uses the velocities update_v! has just written)
    v1 = s.get("v")
    s.apply(ops.shtc_update_rho("wendland2", h, c["dt"], c["m"]))
    xv = np.sum(d * (v1[:, None, :] - v1[None, :, :]), axis=2)
    want_rho = np.where(typ == 0.0, rho + np.sum(dtm * rD * xv, axis=1), rho)
    assert np.max(np.abs(s.get("rho") - want_rho)) <= 1e-13 * np.max(np.abs(want_rho))
    # convect_A!  ldc.jl:96-100: sequential in the visiting order, each pair sees the A the previous ones left
    rho1 = s.get("rho")
    off, ids = s.neighbour_lists()
    want_A = A.copy()
    for p in range(n):
        if typ[p] == c["LID"]:
            continue
        Ap = want_A[p]
        for q in ids[off[p]:off[p + 1]] - 1:
            M = np.outer(v1[p] - v1[q], x[p] - x[q])
            Ap = Ap + (dtm / rho1[p] * float(rDwendland2(h, r[p, q]))) * Ap @ M
        want_A[p] = Ap
    s.apply(ops.shtc_convect_A("wendland2", h, c["dt"], c["m"], c["LID"]))
    got_A = mat(s.get("A"))
    assert np.max(np.abs(got_A - want_A)) <= 1e-13
    assert np.array_equal(got_A[typ == c["LID"]], A[typ == c["LID"]])
    # relax_A!  ldc.jl:102-116 (RK4)
    s.apply(ops.shtc_relax_A(c["dt"], c["tau"]))

    def f(B):
        return -3.0 / c["tau"] * B @ dev(B.transpose(0, 2, 1) @ B)

    A0, dt = got_A, c["dt"]
    k1 = f(A0)
    k2 = f(A0 + dt * k1 / 2)
    k3 = f(A0 + dt * k2 / 2)
    k4 = f(A0 + dt * k3)
    want = A0 + dt * k1 / 6 + dt * k2 / 3 + dt * k3 / 3 + dt * k4 / 6
    assert np.max(np.abs(mat(s.get("A")) - want)) <= 1e-13
    # move!  ldc.jl:129-133
    s.apply(ops.shtc_move(c["dt"]))
    assert np.array_equal(s.get("x"), np.where((typ == 0.0)[:, None], x + s.get("v") * c["dt"], x))


def test_shtc_ldc_time_loop():
    case = configs.shtc_ldc()
    c = case.consts
    s = case.make(OracleSystem)
    for _ in range(150):
        case.step(s)
    assert len(s) == case.n
    v, typ, rho, A = s.get("v"), s.get("type"), s.get("rho"), mat(s.get("A"))
    assert np.all(np.isfinite(v)) and np.all(np.isfinite(A))
    # the lid keeps its prescribed velocity and unit distortion; walls stay at rest; the fluid below the lid is dragged along
    lid = typ == c["LID"]
    assert np.all(v[lid, 0] == c["vlid"]) and np.all(v[typ == 1.0] == 0.0)
    assert np.array_equal(A[lid], np.tile(np.eye(3), (lid.sum(), 1, 1)))
    x = s.get("x")
    top = (typ == 0.0) & (x[:, 1] > 0.97) & (x[:, 0] > 0.3) & (x[:, 0] < 0.7)
    assert np.mean(v[top, 0]) > 0.05
    assert 0.98 < rho.min() and rho.max() < 1.03
    # a plane flow never couples the in-plane block of A to z (A33 itself relaxes through the trace in dev)
    assert np.all(A[:, 2, :2] == 0.0) and np.all(A[:, :2, 2] == 0.0)
