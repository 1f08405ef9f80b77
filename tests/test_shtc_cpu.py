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


# ----------------------------------------------------------------------------- SHTC/beryllium.jl
def wendland2(h, r):
    x = r / h
    return np.where(x <= 1.0, 2.228169203286535 * (1 - x) ** 4 * (1 + 4 * x) / h ** 2, 0.0)


def wendland2h(h, r):  # beryllium.jl:44-47
    x = r / h
    return np.where(x < 1.0, 14.0 * (1.0 - x) ** 3 * (14.0 * x ** 2 - 3.0 * x - 1.0) / (np.pi * h ** 2), 0.0)


def rDwendland2h(h, r):  # beryllium.jl:49-52
    x = r / h
    return np.where(x < 1.0, 140.0 * (1.0 - x) ** 2 * (4.0 - 7.0 * x) / (np.pi * h ** 4), 0.0)


def beryllium_patch(seed=9):
    """One end of the plate, deformed and with perturbed masses, after the script's J0/K0 calibration."""
    case = configs.shtc_beryllium()
    keep = np.flatnonzero(case.init["x"][:, 0] < -0.02)
    rng = np.random.default_rng(seed)
    n = len(keep)
    X = case.init["x"][keep]
    x = X.copy()
    x[:, 1] += 0.8 * (X[:, 0] + 0.03) ** 2 + 0.03 * X[:, 0]
    x[:, 0] += 0.02 * X[:, 1]
    x[:, :2] += rng.uniform(-0.02, 0.02, (n, 2)) * case.consts["dr"]
    A = np.tile(np.eye(3), (n, 1, 1))
    A[:, :2, :2] += rng.uniform(-0.03, 0.03, (n, 2, 2))
    s = OracleSystem(case.fields, case.domain, case.h)
    s.add_particles(x=x, v=rng.uniform(-30, 30, (n, 3)) * np.array([1, 1, 0]), m=case.consts["m0"] * rng.uniform(0.9, 1.1, n),
                    A=A.transpose(0, 2, 1).reshape(n, 9), J0=rng.uniform(-0.02, 0.02, n), K0=rng.uniform(-1e-3, 1e-3, n))
    s.create_cell_list()
    assert len(s) == n
    return case, s


def test_beryllium_operators_against_numpy():
    case, s = beryllium_patch()
    c = case.consts
    h, rho0 = c["h"], c["rho0"]
    x, v, m, A = s.get("x"), s.get("v"), s.get("m"), mat(s.get("A"))
    n = len(x)
    d = x[:, None, :] - x[None, :, :]
    r = np.sqrt(np.sum(d * d, axis=2))
    nb = (r <= h) & ~np.eye(n, dtype=bool)
    ker = np.where(nb, m[None, :] / rho0 * rDwendland2(h, r), 0.0)
    d2 = d[:, :, :2]
    # reset! then find_L!  beryllium.jl:177-184, 140-146
    s.apply(ops.be_reset())
    assert np.array_equal(s.get("J"), s.get("J0")) and np.array_equal(s.get("K"), s.get("K0"))
    s.apply(ops.be_find_L("wendland2", h, rho0))
    T0 = np.einsum("pq,pqi,pqj->pij", ker, d2, d2)
    L0 = np.einsum("pq,pqi,pqj->pij", ker, (v[:, None, :] - v[None, :, :])[:, :, :2], d2)
    assert np.max(np.abs(mat(s.get("T"))[:, :2, :2] - T0)) <= 1e-12 * np.max(np.abs(T0))
    assert np.max(np.abs(mat(s.get("L"))[:, :2, :2] - L0)) <= 1e-12 * np.max(np.abs(L0))
    # update_A!  :148-151
    hdt = 0.5 * c["dt"]
    s.apply(ops.be_update_A(hdt))
    L = L0 @ np.linalg.inv(T0)
    I2 = np.eye(2)
    A2 = A[:, :2, :2] @ (I2 - hdt * L) @ np.linalg.inv(I2 + hdt * L)
    got = mat(s.get("A"))
    assert np.max(np.abs(mat(s.get("L"))[:, :2, :2] - L)) <= 1e-10 * np.max(np.abs(L))
    assert np.max(np.abs(got[:, :2, :2] - A2)) <= 1e-13 and np.all(got[:, 2, 2] == 1.0)
    assert np.all(got[:, 2, :2] == 0.0) and np.all(got[:, :2, 2] == 0.0)
    # reset!, find_J!  :153-158
    s.apply(ops.be_reset())
    s.apply(ops.be_find_J("wendland2", h, rho0))
    mr = np.where(nb, m[None, :] / rho0, 0.0)
    J = s.get("J0") + np.sum(mr * wendland2(h, r), axis=1)
    Kf = s.get("K0") + np.sum(mr * wendland2h(h, r), axis=1)
    np.testing.assert_allclose(s.get("J"), J, rtol=1e-13)
    np.testing.assert_allclose(s.get("K"), Kf, rtol=1e-11, atol=1e-15)
    assert np.max(np.abs(mat(s.get("T"))[:, :2, :2] - T0)) <= 1e-12 * np.max(np.abs(T0))
    # find_T!  :160-164
    s.apply(ops.be_find_T(rho0, c["c_0"], c["c_s"]))
    Afull = got
    G = Afull.transpose(0, 2, 1) @ Afull
    P = 0.5 * rho0 * c["c_0"] ** 2 * ((1.0 - 1.0 / J) / J ** 2 + np.log(J) / J)
    invT = np.zeros((n, 3, 3))
    invT[:, :2, :2] = np.linalg.inv(T0)
    invT[:, 2, 2] = 1.0                                              # the script's 2-D inv, :91-98
    T = (P / rho0)[:, None, None] * np.eye(3) - c["c_s"] ** 2 * G @ dev(G) @ invT
    np.testing.assert_allclose(s.get("P"), P, rtol=1e-10, atol=1e-6 * np.max(np.abs(P)))
    assert np.max(np.abs(mat(s.get("T")) - T)) <= 1e-9 * np.max(np.abs(T))
    # find_f!  :166-175
    s.apply(ops.be_find_f("wendland2", h, rho0, c["c_p"]))
    Tg = mat(s.get("T"))[:, :2, :2]
    Kg = s.get("K")
    kerh = np.where(nb, m[None, :] / rho0 * rDwendland2h(h, r), 0.0)
    f = (-(m[:, None] * ker)[:, :, None] * (np.einsum("pij,pqj->pqi", Tg, d2) + np.einsum("qij,pqj->pqi", Tg, d2))
         - (m[:, None] * kerh * c["c_p"] ** 2 * (Kg[:, None] + Kg[None, :]))[:, :, None] * d2)
    want = np.sum(f, axis=1)
    gotf = s.get("f")
    assert np.max(np.abs(gotf[:, :2] - want)) <= 1e-10 * np.max(np.abs(want)) and np.all(gotf[:, 2] == 0.0)
    # update_v!  :132-134
    s.apply(ops.be_update_v(hdt))
    assert np.array_equal(s.get("v"), v + hdt * gotf / m[:, None])


def test_beryllium_calibration_and_energy_conservation():
    case = configs.shtc_beryllium()
    c = case.consts
    s = case.make(OracleSystem)
    case.prologue(s)
    # the J0/K0 calibration (:117-120) makes the undeformed plate stress-free: J = 1, K = 0, no force
    assert np.max(np.abs(s.get("J") - 1.0)) < 1e-14 and np.max(np.abs(s.get("K"))) < 1e-14
    assert np.max(np.abs(s.get("f"))) < 1e-9 * c["m0"] * c["c_s"] ** 2 / c["h"]
    E0 = configs.beryllium_energy(s, c)
    x0 = s.get("x").copy()
    for _ in range(300):
        case.step(s)
    assert len(s) == case.n
    E1 = configs.beryllium_energy(s, c)
    assert abs(E1 - E0) < 1e-5 * E0                        # measured 7e-7: the symplectic splitting conserves energy
    assert np.max(np.abs(s.get("x") - x0)) > 1e-4          # and the plate does move


# ----------------------------------------------------------------------------- SHTC/twist3d.jl
def rDwendland3(h, r):  # kernels.jl:188-195
    x = r / h
    return np.where(x <= 1.0, -66.84507609859604 * (1 - x) ** 3 / h ** 5, 0.0)


def wendland3(h, r):  # kernels.jl:156-163
    x = r / h
    return np.where(x <= 1.0, 3.3422538049298023 * (1 - x) ** 4 * (1 + 4 * x) / h ** 3, 0.0)


def wendland3h(h, r):  # twist3d.jl:43-46
    x = r / h
    return np.where(x < 1.0, 21.0 * (1.0 - x) ** 3 * (14.0 * x ** 2 - 3.0 * x - 1.0) / (np.pi * h ** 3), 0.0)


def rDwendland3h(h, r):  # twist3d.jl:48-51
    x = r / h
    return np.where(x < 1.0, 210.0 * (1.0 - x) ** 2 * (4.0 - 7.0 * x) / (np.pi * h ** 5), 0.0)


def test_twist3d_operators_against_numpy():
    case = configs.shtc_twist3d(dr=1 / 6)
    c = case.consts
    rng = np.random.default_rng(12)
    keep = np.flatnonzero(case.init["x"][:, 2] < 1.2)               # the clamped foot and the first layers above it
    n = len(keep)
    X = case.init["x"][keep]
    x = X + rng.uniform(-0.05, 0.05, (n, 3)) * c["dr"]
    x[:, 0] += 0.05 * X[:, 2] * X[:, 1]
    x[:, 1] -= 0.05 * X[:, 2] * X[:, 0]
    v = rng.uniform(-20, 20, (n, 3))
    m = c["m0"] * rng.uniform(0.9, 1.1, n)
    A = np.tile(np.eye(3), (n, 1, 1)) + rng.uniform(-0.03, 0.03, (n, 3, 3))
    s = OracleSystem(case.fields, case.domain, case.h)
    s.add_particles(x=x, v=v, m=m, A=A.transpose(0, 2, 1).reshape(n, 9), J0=rng.uniform(-0.02, 0.02, n),
                    K0=rng.uniform(-1e-3, 1e-3, n))
    s.create_cell_list()
    assert len(s) == n
    h, rho0, hdt = c["h"], c["rho0"], 0.5 * c["dt"]
    d = x[:, None, :] - x[None, :, :]
    r = np.sqrt(np.sum(d * d, axis=2))
    nb = (r <= h) & ~np.eye(n, dtype=bool)
    ker = np.where(nb, m[None, :] / rho0 * rDwendland3(h, r), 0.0)
    s.apply(ops.be_reset())
    s.apply(ops.tw_find_L("wendland3", h, rho0))
    T0 = np.einsum("pq,pqi,pqj->pij", ker, d, d)
    L0 = np.einsum("pq,pqi,pqj->pij", ker, v[:, None, :] - v[None, :, :], d)
    assert np.max(np.abs(mat(s.get("T")) - T0)) <= 1e-12 * np.max(np.abs(T0))
    assert np.max(np.abs(mat(s.get("L")) - L0)) <= 1e-12 * np.max(np.abs(L0))
    s.apply(ops.tw_update_A(hdt))
    L = L0 @ np.linalg.inv(T0)
    I3 = np.eye(3)
    A1 = A @ (I3 - hdt * L) @ np.linalg.inv(I3 + hdt * L)
    assert np.max(np.abs(mat(s.get("L")) - L)) <= 1e-9 * np.max(np.abs(L))
    assert np.max(np.abs(mat(s.get("A")) - A1)) <= 1e-12
    s.apply(ops.be_reset())
    s.apply(ops.tw_find_J("wendland3", h, rho0))
    mr = np.where(nb, m[None, :] / rho0, 0.0)
    J = s.get("J0") + np.sum(mr * wendland3(h, r), axis=1)
    Kf = s.get("K0") + np.sum(mr * wendland3h(h, r), axis=1)
    np.testing.assert_allclose(s.get("J"), J, rtol=1e-13)
    np.testing.assert_allclose(s.get("K"), Kf, rtol=1e-10, atol=1e-14)
    s.apply(ops.tw_find_T(rho0, c["c_0"], c["c_s"]))
    Fm = np.linalg.inv(A1)
    B = Fm @ Fm.transpose(0, 2, 1)
    detF = 1.0 / J
    P = -rho0 * c["c_0"] ** 2 * detF ** 2 * (detF - 1.0)
    T = (-P / rho0)[:, None, None] * I3 - c["c_s"] ** 2 * (B - I3) @ np.linalg.inv(T0)
    np.testing.assert_allclose(s.get("P"), P, rtol=1e-10, atol=1e-9 * np.max(np.abs(P)))
    assert np.max(np.abs(mat(s.get("T")) - T)) <= 1e-9 * np.max(np.abs(T))
    s.apply(ops.tw_find_f("wendland3", h, rho0, c["c_p"]))
    Tg, Kg = mat(s.get("T")), s.get("K")
    kerh = np.where(nb, m[None, :] / rho0 * rDwendland3h(h, r), 0.0)
    f = ((m[:, None] * ker)[:, :, None] * (np.einsum("pij,pqj->pqi", Tg, d) + np.einsum("qij,pqj->pqi", Tg, d))
         - (m[:, None] * kerh * c["c_p"] ** 2 * (Kg[:, None] + Kg[None, :]))[:, :, None] * d)
    want = np.sum(f, axis=1)
    assert np.max(np.abs(s.get("f") - want)) <= 1e-10 * np.max(np.abs(want))
    v0, fg = s.get("v"), s.get("f")
    s.apply(ops.tw_update_v(hdt))
    assert np.array_equal(s.get("v"), np.where((x[:, 2] > 0.0)[:, None], v0 + hdt * fg / m[:, None], v0))


def test_twist3d_time_loop_on_the_oracle():
    case = configs.shtc_twist3d(dr=1 / 8)
    c = case.consts
    s = case.make(OracleSystem)
    case.prologue(s)
    assert np.max(np.abs(s.get("J") - 1.0)) < 1e-13 and np.max(np.abs(s.get("K"))) < 1e-13     # calibration :110-113
    x0 = s.get("x").copy()
    for _ in range(60):
        case.step(s)
    assert len(s) == case.n
    x, A = s.get("x"), mat(s.get("A"))
    assert np.all(np.isfinite(x)) and np.all(np.isfinite(A))
    foot = x0[:, 2] <= 0.0
    assert np.array_equal(x[foot], x0[foot])                           # the clamped foot (:125-129) never moves
    top = x0[:, 2] > 0.9 * c["H"]
    ang0, ang1 = np.arctan2(x0[top, 1], x0[top, 0]), np.arctan2(x[top, 1], x[top, 0])
    rad = np.hypot(x0[top, 0], x0[top, 1]) > 0.2
    turn = np.angle(np.exp(1j * (ang1 - ang0)))[rad]
    assert np.all(turn < 0.0) and np.mean(turn) < -0.05                # the top spins clockwise (init_velocity :36-38)
    dets = np.linalg.det(A)
    assert 0.9 < dets.min() and dets.max() < 1.1                       # nearly incompressible rubber (nu = 0.495)


# ----------------------------------------------------------------------------- SHTC/taco.jl
def test_taco_operators_against_numpy():
    case = configs.shtc_taco()
    c = case.consts
    rng = np.random.default_rng(21)
    keep = np.flatnonzero((case.init["x"][:, 0] > 0.6) & (np.abs(case.init["x"][:, 1]) < 0.35))   # a sector with both walls
    n = len(keep)
    x = case.init["x"][keep] + rng.uniform(-0.05, 0.05, (n, 3)) * c["dr"] * np.array([1, 1, 0])
    typ = case.init["type"][keep]
    assert set(np.unique(typ)) == {0.0, 1.0, 2.0}
    m = c["m0"] * rng.uniform(0.9, 1.1, n)
    A = np.tile(np.eye(3), (n, 1, 1))
    A[:, :2, :2] += rng.uniform(-0.03, 0.03, (n, 2, 2))
    A[:, 2, 2] += rng.uniform(-0.01, 0.01, n)
    s = OracleSystem(case.fields, case.domain, case.h)
    s.add_particles(x=x, x0=case.init["x"][keep], v=rng.uniform(-1, 1, (n, 3)) * np.array([1, 1, 0]), m=m, type=typ,
                    A=A.transpose(0, 2, 1).reshape(n, 9), C_rho=rng.uniform(-0.02, 0.02, n), C_lambda=rng.uniform(-0.5, 0.5, n))
    s.create_cell_list()
    assert len(s) == n
    h = c["h"]
    d = x[:, None, :] - x[None, :, :]
    r = np.sqrt(np.sum(d * d, axis=2))
    nb = (r <= h) & ~np.eye(n, dtype=bool)
    d2 = d[:, :, :2]
    # reset! + find_rho! with self = true  (taco.jl:164-170, 141-146, 252)
    s.apply(ops.be_reset(J="rho", Kf="lambda", J0="C_rho", K0="C_lambda"))
    s.apply(ops.be_find_J("wendland2", h, 1.0, J="rho", Kf="lambda"), self_=True)
    mq = np.where(nb, m[None, :], 0.0)
    rho = s.get("C_rho") + np.sum(mq * wendland2(h, r), axis=1) + m * float(wendland2(h, 0.0))
    lam = s.get("C_lambda") + np.sum(mq * wendland2h(h, r), axis=1) + m * float(wendland2h(h, 0.0))
    T0 = np.einsum("pq,pqi,pqj->pij", mq * rDwendland2(h, r), d2, d2)
    np.testing.assert_allclose(s.get("rho"), rho, rtol=1e-13)
    np.testing.assert_allclose(s.get("lambda"), lam, rtol=1e-11, atol=1e-13)
    assert np.max(np.abs(mat(s.get("T"))[:, :2, :2] - T0)) <= 1e-12 * np.max(np.abs(T0))
    # find_T!  :148-152
    s.apply(ops.ta_find_T(c["rho0"], c["c_0"], c["c_s"]))
    G = A.transpose(0, 2, 1) @ A
    P = c["c_0"] ** 2 * (rho - c["rho0"]) * c["rho0"] / rho
    si = np.zeros((n, 3, 3))
    si[:, :2, :2] = np.linalg.inv(T0)                                # subinv, tools.jl:45-52
    T = (-P / rho ** 2)[:, None, None] * np.eye(3) + c["c_s"] ** 2 * G @ dev(G) @ si
    np.testing.assert_allclose(s.get("P"), P, rtol=1e-10, atol=1e-12)
    assert np.max(np.abs(mat(s.get("T")) - T)) <= 1e-9 * np.max(np.abs(T))
    # find_f!  :154-162
    s.apply(ops.ta_find_f("wendland2", h, c["c_p"], c["rho0"]))
    Tg, lg = mat(s.get("T")), s.get("lambda")
    ker = mq * rDwendland2(h, r)
    kerh = mq * rDwendland2h(h, r)
    f = ((m[:, None] * ker)[:, :, None] * np.einsum("pqij,pqj->pqi", Tg[:, None] + Tg[None, :], d)
         - (m[:, None] * kerh * (c["c_p"] / c["rho0"]) ** 2 * (lg[:, None] + lg[None, :]))[:, :, None] * d)
    want = np.sum(f, axis=1)
    assert np.max(np.abs(s.get("f") - want)) <= 1e-10 * np.max(np.abs(want))
    # update_v!  :108-114 and update_x!  :116-126
    v0, fg, x1 = s.get("v"), s.get("f"), s.get("x")
    hdt = 0.5 * c["dt"]
    s.apply(ops.ta_update_v(hdt, c["R1"], c["R2"], c["omega"]))
    want_v = np.where((typ == 0.0)[:, None], v0 + hdt * fg / m[:, None], configs.taco_exact_velocity(x1, c))
    np.testing.assert_allclose(s.get("v"), want_v, rtol=1e-15, atol=1e-18)
    t = 0.37
    s.apply(ops.ta_update_x(hdt, c["omega"], t, c["OUTER"]))
    X0 = s.get("x0")
    cw, sw = np.cos(c["omega"] * t), np.sin(c["omega"] * t)
    rot = np.column_stack([X0[:, 0] * cw - X0[:, 1] * sw, X0[:, 0] * sw + X0[:, 1] * cw, np.zeros(n)])
    want_x = np.where((typ == 0.0)[:, None], x1 + hdt * s.get("v"), np.where((typ == 2.0)[:, None], rot, x1))
    np.testing.assert_allclose(s.get("x"), want_x, rtol=1e-15, atol=1e-18)


def test_taco_calibration_and_spin_up():
    case = configs.shtc_taco()
    c = case.consts
    s = case.make(OracleSystem)
    case.prologue(s)
    # C_rho / C_lambda (:94-97) make the initial state uniform: rho = rho0, lambda = 0, no force
    assert np.max(np.abs(s.get("rho") - c["rho0"])) < 1e-14 and np.max(np.abs(s.get("lambda"))) < 1e-13
    assert np.max(np.abs(s.get("f"))) < 1e-12
    for _ in range(200):
        case.step(s)
    assert len(s) == case.n
    x, v, typ = s.get("x"), s.get("v"), s.get("type")
    r = np.hypot(x[:, 0], x[:, 1])
    outer = typ == c["OUTER"]
    # the outer cylinder turns rigidly: radius kept, angle advanced by omega*t
    r0 = np.hypot(case.init["x"][outer, 0], case.init["x"][outer, 1])
    assert np.max(np.abs(r[outer] - r0)) < 1e-12
    assert np.allclose(v[~(typ == 0.0)], configs.taco_exact_velocity(x, c)[~(typ == 0.0)], rtol=1e-12, atol=1e-14)
    # and drags the fluid next to it along (anticlockwise), while the fluid at the resting inner cylinder stays slow
    vphi = (-x[:, 1] * v[:, 0] + x[:, 0] * v[:, 1]) / r
    near_outer = (typ == 0.0) & (r > 1.93)
    near_inner = (typ == 0.0) & (r < 1.1)
    assert np.mean(vphi[near_outer]) > 0.05 and np.mean(vphi[near_outer]) > 3 * abs(np.mean(vphi[near_inner]))
    assert np.all(np.isfinite(s.get("A")))
