"""Registered operators: the named, parameterised counterparts of the closures the reference's examples
pass to ``apply!``.  Each factory returns an :class:`Operator` carrying the
operator id, the field binding (names of the particle fields playing each role) and the Float64 parameter
block.  Parameter expressions are folded here exactly as the reference's ``const`` expressions are
(e.g. ``2*nu``, ``c^2``, ``0.5*dt``), so host and device see the same doubles.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

from . import abi

K = abi.K


@dataclass(frozen=True)
class Operator:
    op: int
    fields: Tuple[str, ...]
    params: Tuple[float, ...]
    binary: bool
    name: str = ""


def _kid(kernel) -> float:
    return float(abi.KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel))


# ---- WCSPH: examples/collapse_dry.jl, collapse3d.jl, cavity_flow.jl
def balance_of_mass(kernel, m, h, nu=0.0, x="x", v="v", rho="rho", Drho="Drho"):
    """collapse_dry.jl:112-115 / collapse3d.jl:87-90; cavity_flow.jl:92-94 with nu = 0."""
    return Operator(K["SP_OP_BALANCE_OF_MASS"], (x, v, rho, Drho), (_kid(kernel), m, h, 2 * nu), True,
                    "balance_of_mass!")


def find_pressure(dt, c, rho0, P0=0.0, rho="rho", Drho="Drho", P="P"):
    """collapse_dry.jl:123-127; cavity_flow.jl:96-100 adds P0."""
    return Operator(K["SP_OP_FIND_PRESSURE"], (rho, Drho, P), (dt, c * c, rho0, P0), False, "find_pressure!")


def internal_force(kernel, m, h, mu, rho0, x="x", v="v", P="P", rho="rho", Dv="Dv", type="type"):
    """collapse_dry.jl:135-141."""
    return Operator(K["SP_OP_INTERNAL_FORCE"], (x, v, P, rho, Dv, type), (_kid(kernel), m, h, mu, rho0), True,
                    "internal_force!")


def internal_force_cavity(m, h, Re, vlid, ylid=1.0, lid_type=2.0, x="x", v="v", P="P", rho="rho", Dv="Dv",
                          type="type"):
    """cavity_flow.jl:102-114."""
    return Operator(K["SP_OP_INTERNAL_FORCE_CAVITY"], (x, v, P, rho, Dv, type),
                    (m, h, float(Re), vlid, ylid, lid_type), True, "internal_force! (cavity)")


def move(dtm, x="x", v="v", Dv="Dv", type="type"):
    """collapse_dry.jl:148-153 (dtm = 0.5*dt), collapse3d.jl:106-111 (dtm = dt)."""
    return Operator(K["SP_OP_MOVE"], (x, v, Dv, type), (dtm,), False, "move!")


def accelerate(hdt, g=(0.0, 0.0, 0.0), v="v", Dv="Dv", type="type"):
    """collapse_dry.jl:155-159: v += hdt*(Dv + g) for fluid particles."""
    return Operator(K["SP_OP_ACCELERATE"], (v, Dv, type), (hdt, g[0], g[1], g[2]), False, "accelerate!")


# ---- tests/test_collision_2d.jl
def density_sum(kernel, m, h, out="rho", x="x"):
    """find_rho! / find_rho0!, test_collision_2d.jl:63-69 (use with self=True)."""
    return Operator(K["SP_OP_DENSITY_SUM"], (x, out), (_kid(kernel), m, h), True, "find_rho!")


def pressure_from_rho(c, rho="rho", rho0="rho0", P="P"):
    return Operator(K["SP_OP_PRESSURE_FROM_RHO"], (rho, rho0, P), (c * c,), False, "find_pressure! (collision)")


def internal_force_sym(kernel, m, h, rho0, x="x", P="P", a="a"):
    return Operator(K["SP_OP_INTERNAL_FORCE_SYM"], (x, P, a), (_kid(kernel), m, h, rho0), True,
                    "internal_force! (collision)")


def fill(field, value=0.0):
    return Operator(K["SP_OP_FILL"], (field,), (value,), False, "reset!")


def advect(dt, x="x", v="v"):
    return Operator(K["SP_OP_ADVECT"], (x, v), (dt,), False, "move! (collision)")


def kick(hdt, v="v", a="a"):
    return Operator(K["SP_OP_KICK"], (v, a), (hdt,), False, "accelerate! (collision)")


# ---- examples/drop.jl
def find_normal(kernel, vol, h, x="x", n="n"):
    """drop.jl:76-78: n_p += 2*vol*vol*rDw(h,r)*x_pq."""
    return Operator(K["SP_OP_FIND_NORMAL"], (x, n), (_kid(kernel), 2 * vol * vol, h), True, "find_n!")


def normalize(s0, n="n"):
    """drop.jl:84-87."""
    return Operator(K["SP_OP_NORMALIZE"], (n,), (s0,), False, "normalize_n!")


def internal_force_tension(m, h, mu, rho0, beta, s0, x="x", v="v", P="P", n="n", a="a"):
    """drop.jl:101-113 (rDwendland3 / DDwendland3)."""
    return Operator(K["SP_OP_INTERNAL_FORCE_TENSION"], (x, v, P, n, a), (m, h, mu, rho0, beta, s0), True,
                    "internal_force! (surface tension)")


# ---- examples/collapse_symplectic.jl, examples/Kepler_vortex.jl (reversible fixed-point integrator, LJ walls)
def density_sum_fluid(kernel, m, h, out="rho", x="x", type="type"):
    """find_rho! / find_rho0!, collapse_symplectic.jl:98-108: fluid-fluid pairs only (use with self=True)."""
    return Operator(K["SP_OP_DENSITY_SUM_FLUID"], (x, out, type), (_kid(kernel), m, h), True, "find_rho! (fluid)")


def internal_force_lj(kernel, m, h, dr_wall, E_wall, eps, rho0=0.0, wall_type=1.0, x="x", P="P", rho="rho", a="a",
                      type="type"):
    """collapse_symplectic.jl:114-123 (rho0 = 0: P/rho^2 of each particle) and Kepler_vortex.jl:155-164 (P/rho0^2):
    pressure force between fluid particles, Lennard-Jones repulsion from wall particles closer than dr_wall."""
    return Operator(K["SP_OP_INTERNAL_FORCE_LJ"], (x, P, rho, a, type),
                    (_kid(kernel), m, h, rho0, wall_type, dr_wall, E_wall, eps), True, "internal_force! (LJ walls)")


def move_rev(dt, x="x", v="v", type="type"):
    """collapse_symplectic.jl:134-138: x = rev_add(x, dt*v) for fluid particles (utils/FixPA.jl)."""
    return Operator(K["SP_OP_MOVE_REV"], (x, v, type), (dt,), False, "move! (rev_add)")


def accelerate_rev(hdt, g=(0.0, 0.0, 0.0), v="v", a="a", type="type"):
    """collapse_symplectic.jl:140-144: v = rev_add(v, hdt*(a + g)) for fluid particles."""
    return Operator(K["SP_OP_ACCELERATE_REV"], (v, a, type), (hdt, g[0], g[1], g[2]), False, "accelerate! (rev_add)")


def accelerate_rev_central(hdt, GM, x="x", v="v", a="a", type="type"):
    """Kepler_vortex.jl:180-184: v = rev_add(v, hdt*rev_add(a, -GM/norm(x)^3*x)) for fluid particles."""
    return Operator(K["SP_OP_ACCELERATE_REV_CENTRAL"], (x, v, a, type), (hdt, GM), False,
                    "accelerate! (rev_add, central gravity)")


def lj_potential(h, m, E_wall, dr_wall, eps, wall_type=1.0, out="U", x="x", type="type"):
    """sum(sys, LJ_potential, p) for every particle p (core.jl:271-291, collapse_symplectic.jl:146-153):
    out_p += m*E_wall*(0.5 s^2 - 0.25 s^4 - 0.25), s = dr_wall/(r + eps), over wall neighbours with r < dr_wall."""
    return Operator(K["SP_OP_LJ_POTENTIAL"], (x, out, type), (h, m * E_wall, wall_type, dr_wall, eps), True,
                    "LJ_potential")


# ---- examples/cylinder.jl (per-particle mass, inflow buffer)
def cyl_balance_of_mass(kernel, h, nu, x="x", v="v", rho="rho", Drho="Drho", m="m", type="type"):
    """cylinder.jl:102-108: ker = q.m*rDw; Drho += ker*dot(x_pq, v_pq); fluid-fluid pairs add 2*nu/rho_p*(rho_p - rho_q)."""
    return Operator(K["SP_OP_CYL_BALANCE_OF_MASS"], (x, v, rho, Drho, m, type), (_kid(kernel), h, 2 * nu), True,
                    "balance_of_mass! (cylinder)")


def cyl_find_pressure(dt, c, rho0, x1_min, x="x", rho="rho", Drho="Drho", P="P"):
    """cylinder.jl:110-116: the density is integrated only downstream of x1_min = -bc_width + h."""
    return Operator(K["SP_OP_CYL_FIND_PRESSURE"], (x, rho, Drho, P), (dt, c * c, rho0, x1_min), False,
                    "find_pressure! (cylinder)")


def cyl_internal_force(kernel, h, mu, x="x", v="v", P="P", rho="rho", a="a", m="m"):
    """cylinder.jl:118-123: pressure and Monaghan viscosity with the neighbour's mass."""
    return Operator(K["SP_OP_CYL_INTERNAL_FORCE"], (x, v, P, rho, a, m), (_kid(kernel), h, mu, 0.01 * h * h), True,
                    "internal_force! (cylinder)")


def move_types(dt, type_a, type_b, x="x", v="v", a="a", type="type"):
    """cylinder.jl:125-130: a = 0; particles of the two given types move."""
    return Operator(K["SP_OP_MOVE_TYPES"], (x, v, a, type), (dt, type_a, type_b), False, "move! (cylinder)")


def cyl_accelerate(hdt, cyl1, U_max, x="x", v="v", a="a", type="type"):
    """cylinder.jl:132-143: v += hdt*(a + gravity(p)), gravity = 0.3*U_max^2*f/|f|^2 towards (cyl1, 0)."""
    return Operator(K["SP_OP_CYL_ACCELERATE"], (x, v, a, type), (hdt, cyl1, 0.3 * U_max ** 2), False,
                    "accelerate! (cylinder)")


def set_inflow_speed(t, t_acc, U_max, chan_w, inflow_type=1.0, x="x", v="v", type="type"):
    """cylinder.jl:91-97 at time t: parabolic inflow profile ramped over t_acc."""
    return Operator(K["SP_OP_SET_INFLOW_SPEED"], (x, v, type), (inflow_type, min(1.0, t / t_acc), U_max, chan_w), False,
                    "set_inflow_speed!")


# ---- examples/rod.jl (elastic solid, tensor-valued fields A, H, B with 9 components each)
def rod_find_A(kernel, h, x="x", X="X", A="A", H="H"):
    """rod.jl:128-134: A += -w*outer(X_pq, x_pq); H += -w*outer(x_pq, x_pq)."""
    return Operator(K["SP_OP_ROD_FIND_A"], (x, X, A, H), (_kid(kernel), h), True, "find_A!")


def rod_find_B(m, c_l, c_s, A="A", H="H", B="B"):
    """rod.jl:136-143: A = A*inv(H); B = m*(P*inv(A') + c_s^2*A*dev(A'A))*inv(H), P = c_l^2*(det A - 1)."""
    return Operator(K["SP_OP_ROD_FIND_B"], (A, H, B), (m, c_l, c_s), False, "find_B!")


def rod_find_f(kernel, h, m, vol, nu, x="x", v="v", X="X", A="A", B="B", f="f"):
    """rod.jl:145-160: elastic force with the energy-conserving "eta" correction and artificial viscosity."""
    return Operator(K["SP_OP_ROD_FIND_F"], (x, v, X, A, B, f), (_kid(kernel), h, 2 * m * vol, nu), True, "find_f!")


def rod_pull(X1_min, fy, X="X", f="f"):
    """rod.jl:162-166: the free end is pulled upwards."""
    return Operator(K["SP_OP_ROD_PULL"], (X, f), (X1_min, fy), False, "pull!")


def rod_update_v(hdt, m, X1_clamp, v="v", f="f", X="X"):
    """rod.jl:168-174: v += hdt*f/m; the clamped end (X[1] < X1_clamp) stays at rest."""
    return Operator(K["SP_OP_ROD_UPDATE_V"], (v, f, X), (hdt, m, X1_clamp), False, "update_v!")


def rod_update_x(dt, x="x", v="v", A="A", H="H", f="f", e="e"):
    """rod.jl:176-183: x += dt*v and the per-step accumulators are reset."""
    return Operator(K["SP_OP_ROD_UPDATE_X"], (x, v, A, H, f, e), (dt,), False, "update_x!")


def rod_find_e(h, x="x", X="X", A="A", e="e"):
    """rod.jl:185-188: e += |inv(A_p)*X_pq - x_pq|^2."""
    return Operator(K["SP_OP_ROD_FIND_E"], (x, X, A, e), (h,), True, "find_e!")


# ---- examples/SHTC/ldc.jl (SHTC fluid: full 3x3 distortion field A and stress tensor, 9-component fields)
def shtc_find_stress(c_l, c_s, rho0, acf, A="A", rho="rho", stress="stress"):
    """ldc.jl:118-121: stress = c_l^2*(rho - rho0/(1 + acf))*I + c_s^2*rho*G*dev(G), G = A'*A."""
    return Operator(K["SP_OP_SHTC_FIND_STRESS"], (A, rho, stress), (c_l, c_s, rho0 / (1.0 + acf)), False, "find_stress!")


def shtc_update_v(kernel, h, dt, m, x="x", v="v", rho="rho", stress="stress", type="type"):
    """ldc.jl:123-127."""
    return Operator(K["SP_OP_SHTC_UPDATE_V"], (x, v, rho, stress, type), (_kid(kernel), h, dt * m), True, "update_v!")


def shtc_update_rho(kernel, h, dt, m, x="x", v="v", rho="rho", type="type"):
    """ldc.jl:90-94."""
    return Operator(K["SP_OP_SHTC_UPDATE_RHO"], (x, v, rho, type), (_kid(kernel), h, dt * m), True, "update_rho!")


def shtc_convect_A(kernel, h, dt, m, skip_type, x="x", v="v", rho="rho", A="A", type="type"):
    """ldc.jl:96-100: A_p += dt*m/rho_p*rDw*A_p*(v_pq*x_pq'), pair after pair in the reference's visiting order."""
    return Operator(K["SP_OP_SHTC_CONVECT_A"], (x, v, rho, A, type), (_kid(kernel), h, dt * m, skip_type), True,
                    "convect_A!")


def shtc_relax_A(dt, tau, A="A"):
    """ldc.jl:102-116: one RK4 step of the strain relaxation dA/dt = -3/tau*A*dev(A'*A)."""
    return Operator(K["SP_OP_SHTC_RELAX_A"], (A,), (dt, tau), False, "relax_A!")


def shtc_move(dt, x="x", v="v", type="type"):
    """ldc.jl:129-133."""
    return Operator(K["SP_OP_SHTC_MOVE"], (x, v, type), (dt,), False, "move! (SHTC)")


# ---- examples/SHTC/beryllium.jl (SHTC solid in 2-D; T, L, A are 9-component RealMatrix fields)
def be_find_L(kernel, h, rho0, x="x", v="v", m="m", T="T", L="L"):
    """beryllium.jl:140-146: T += ker*outer(x_pq, x_pq); L += ker*outer(v_pq, x_pq), ker = m_q/rho0*rDw."""
    return Operator(K["SP_OP_BE_FIND_L"], (x, v, m, T, L), (_kid(kernel), h, rho0), True, "find_L!")


def be_update_A(hdt, A="A", T="T", L="L"):
    """beryllium.jl:148-151: L = L*inv(T); A = A*(I - hdt*L)*inv(I + hdt*L)."""
    return Operator(K["SP_OP_BE_UPDATE_A"], (A, T, L), (hdt,), False, "update_A!")


def be_find_J(kernel, h, rho0, x="x", m="m", T="T", J="J", Kf="K"):
    """beryllium.jl:153-158."""
    return Operator(K["SP_OP_BE_FIND_J"], (x, m, T, J, Kf), (_kid(kernel), h, rho0), True, "find_J!")


def be_find_T(rho0, c_0, c_s, A="A", T="T", P="P", J="J"):
    """beryllium.jl:160-164."""
    return Operator(K["SP_OP_BE_FIND_T"], (A, T, P, J), (rho0, c_0, c_s), False, "find_T!")


def be_find_f(kernel, h, rho0, c_p, x="x", m="m", T="T", Kf="K", f="f"):
    """beryllium.jl:166-175: stress force of both particles and the anti-clumping force."""
    return Operator(K["SP_OP_BE_FIND_F"], (x, m, T, Kf, f), (_kid(kernel), h, rho0, c_p), True, "find_f!")


def be_reset(f="f", L="L", T="T", J="J", Kf="K", J0="J0", K0="K0"):
    """beryllium.jl:177-184."""
    return Operator(K["SP_OP_BE_RESET"], (f, L, T, J, Kf, J0, K0), (), False, "reset!")


def be_update_v(hdt, v="v", f="f", m="m"):
    """beryllium.jl:132-134: v += hdt*f/m."""
    return Operator(K["SP_OP_BE_UPDATE_V"], (v, f, m), (hdt,), False, "update_v!")


# ---- examples/SHTC/twist3d.jl (SHTC solid in 3-D; full 3x3 T, L, A)
def tw_find_L(kernel, h, rho0, x="x", v="v", m="m", T="T", L="L"):
    """twist3d.jl:135-141."""
    return Operator(K["SP_OP_TW_FIND_L"], (x, v, m, T, L), (_kid(kernel), h, rho0), True, "find_L! (3-D)")


def tw_update_A(hdt, A="A", T="T", L="L"):
    """twist3d.jl:143-146."""
    return Operator(K["SP_OP_TW_UPDATE_A"], (A, T, L), (hdt,), False, "update_A! (3-D)")


def tw_find_J(kernel, h, rho0, x="x", m="m", T="T", J="J", Kf="K"):
    """twist3d.jl:148-153."""
    return Operator(K["SP_OP_TW_FIND_J"], (x, m, T, J, Kf), (_kid(kernel), h, rho0), True, "find_J! (3-D)")


def tw_find_T(rho0, c_0, c_s, A="A", T="T", P="P", J="J"):
    """twist3d.jl:155-161."""
    return Operator(K["SP_OP_TW_FIND_T"], (A, T, P, J), (rho0, c_0, c_s), False, "find_T! (3-D)")


def tw_find_f(kernel, h, rho0, c_p, x="x", m="m", T="T", Kf="K", f="f"):
    """twist3d.jl:163-172."""
    return Operator(K["SP_OP_TW_FIND_F"], (x, m, T, Kf, f), (_kid(kernel), h, rho0, c_p), True, "find_f! (3-D)")


def tw_update_v(hdt, x="x", v="v", f="f", m="m"):
    """twist3d.jl:125-129: the part of the column below z = 0 is clamped."""
    return Operator(K["SP_OP_TW_UPDATE_V"], (x, v, f, m), (hdt,), False, "update_v! (3-D)")


# ---- examples/SHTC/taco.jl (Taylor-Couette flow, SHTC fluid); find_L!/update_A!/reset!/find_rho! are the be_* operators
# with rho0 = 1, relax_A! is shtc_relax_A
def ta_find_T(rho0, c_0, c_s, A="A", T="T", P="P", rho="rho"):
    """taco.jl:148-152."""
    return Operator(K["SP_OP_TA_FIND_T"], (A, T, P, rho), (rho0, c_0, c_s), False, "find_T! (taco)")


def ta_find_f(kernel, h, c_p, rho0, x="x", m="m", T="T", lam="lambda", f="f"):
    """taco.jl:154-162."""
    return Operator(K["SP_OP_TA_FIND_F"], (x, m, T, lam, f), (_kid(kernel), h, (c_p / rho0) ** 2), True, "find_f! (taco)")


def ta_update_v(hdt, R1, R2, omega, x="x", v="v", f="f", m="m", type="type"):
    """taco.jl:108-114: the fluid is kicked, wall particles carry the exact Couette velocity (vexact :39-42)."""
    return Operator(K["SP_OP_TA_UPDATE_V"], (x, v, f, m, type), (hdt, R1, R2, omega), False, "update_v! (taco)")


def ta_update_x(hdt, omega, t, outer_type, x="x", v="v", x0="x0", type="type"):
    """taco.jl:116-126 at time t: the fluid drifts, the outer cylinder is rotated rigidly from its initial position."""
    import math
    return Operator(K["SP_OP_TA_UPDATE_X"], (x, v, x0, type), (hdt, math.cos(omega * t), math.sin(omega * t), outer_type),
                    False, "update_x! (taco)")


# ---- examples/static_container.jl
def sc_balance_of_mass(kernel, m, h, dt, x="x", v="v", rho="rho"):
    """static_container.jl:102-104: the density is integrated inside the pair loop."""
    return Operator(K["SP_OP_SC_BALANCE_OF_MASS"], (x, v, rho), (_kid(kernel), m, h, dt), True, "balance_of_mass! (sc)")


def sc_internal_force(kernel, m, h, mu, c, rho0, x="x", v="v", rho="rho", a="a", type="type"):
    """static_container.jl:106-114 with pressure(p) = c^2*(rho - rho0) (:68-70)."""
    return Operator(K["SP_OP_SC_INTERNAL_FORCE"], (x, v, rho, a, type), (_kid(kernel), m, h, mu, c * c, rho0), True,
                    "internal_force! (sc)")


def move_all(dtm, x="x", v="v", a="a"):
    """static_container.jl:116-119."""
    return Operator(K["SP_OP_MOVE_ALL"], (x, v, a), (dtm,), False, "move! (all)")


# ---- ISPH: examples/collapse_dry_implicit.jl
def isph_initialize(dt, g, x="x", v="v", div="div", L="L", lam="lambda", type="type"):
    return Operator(K["SP_OP_ISPH_INITIALIZE"], (x, v, div, L, lam, type), (dt, g[0], g[1], g[2]), False,
                    "initialize!")


def isph_viscous_force(kernel, m, h, mu, rho, x="x", v="v", Dv="Dv"):
    return Operator(K["SP_OP_ISPH_VISCOUS_FORCE"], (x, v, Dv), (_kid(kernel), m, h, mu, rho), True, "viscous_force!")


def isph_div_L_lambda(kernel, m, h, rho, dim, x="x", v="v", div="div", L="L", lam="lambda"):
    return Operator(K["SP_OP_ISPH_DIV_L_LAMBDA"], (x, v, div, L, lam), (_kid(kernel), m, h, rho, float(dim)), True,
                    "div_L_lambda!")


def isph_projection_vector(h, dt, div="div", b="b"):
    return Operator(K["SP_OP_ISPH_PROJECTION_VECTOR"], (div, b), (h, dt), False, "projection_vector")


def isph_internal_force(kernel, m, h, rho, x="x", P="P", Dv="Dv"):
    return Operator(K["SP_OP_ISPH_INTERNAL_FORCE"], (x, P, Dv), (_kid(kernel), m, h, rho), True,
                    "internal_force! (isph)")


def isph_accelerate(dt, v="v", Dv="Dv", type="type"):
    return Operator(K["SP_OP_ISPH_ACCELERATE"], (v, Dv, type), (dt,), False, "accelerate! (isph)")


@dataclass(frozen=True)
class PoissonOperator:
    """projection_matrix of collapse_dry_implicit.jl:154-163 as a matrix-free operator."""
    fields: Tuple[str, ...]  # x, L, lambda, type
    params: Tuple[float, ...]  # kernel, m, h, rho, C_free


def isph_projection_matrix(kernel, m, h, rho, C_free, x="x", L="L", lam="lambda", type="type"):
    return PoissonOperator((x, L, lam, type), (_kid(kernel), m, h, rho, C_free))
