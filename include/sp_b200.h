/*
 * sp_b200.h — C ABI of the B200-native neighbour-search / pair-sweep engine.
 *
 * This is the drop-in boundary for the hot path of SmoothedParticles.jl
 * (reference = /root/reference, pure Julia).  Every entry point is `extern "C"`,
 * takes plain pointers and sizes, returns an int32 status (0 = SP_OK) and never
 * throws.  A Julia host binds these with `ccall` (see INTEGRATION.md and
 * julia/SmoothedParticlesB200.jl); the Python host in
 * smoothedparticles.jl_b200/ binds them with ctypes.
 *
 * Reference interface each group replaces (file:line relative to the reference):
 *   sp_create / sp_key_params        ParticleSystem ctor          src/structs.jl:57-91
 *   sp_add_field / sp_upload / ...   struct-of-arrays storage of  src/structs.jl:53 (particles::Vector{T})
 *                                    and ParticleField            src/structs.jl:118-125
 *   sp_create_cell_list              create_cell_list!            src/core.jl:51-90 (+ find_key structs.jl:97-106,
 *                                                                 is_inside(Box) geometry.jl:24-30)
 *   sp_apply                         apply! / apply_unary! /      src/core.jl:94-161
 *                                    apply_binary!
 *   sp_sum_at_points                 SmoothedParticles.sum(sys,f,x)  src/core.jl:240-260
 *   sp_poisson_apply / sp_poisson_cg assemble_matrix + cg         src/core.jl:196-225,
 *                                                                 examples/collapse_dry_implicit.jl:154-163,223-227
 *   sp_reduce                        energy / get_globals loops   examples/collapse_dry.jl:166-187
 *   sp_kernel_eval                   kernel functions             src/kernels.jl
 *   sp_generate_particles            generate_particles!          src/grids.jl:52-144,253-258, src/geometry.jl:15-258
 *   sp_respawn                       add_new_particles! (inflow)  examples/cylinder.jl:145-156
 *   sp_assemble_matrix               assemble_matrix              src/core.jl:196-225
 *   sp_run_program                   the examples' time loops     examples/collapse3d.jl:134-151, collapse_dry.jl:202-211
 *   sp_graph_*                       (no counterpart: a host loop body recorded once, replayed as one CUDA graph launch)
 *   sp_slab_*                        (no counterpart: one process per GPU, slab decomposition over NCCL)
 *
 * Because arbitrary Julia closures cannot run inside CUDA kernels, the per-pair
 * and per-particle actions of the shipped examples are *registered operators*
 * (SP_OP_*), each taking a list of field ids (the "binding": which device field
 * plays x, v, rho, ...) and a block of Float64 parameters that the host computes
 * exactly as the reference's `const` expressions are folded (e.g. 0.5*dt).
 *
 * Threading: a handle is not thread-safe; calls on one handle are ordered on one
 * CUDA stream.  Calls that return host data synchronise; the others may return
 * before the device finished (errors then surface at the next synchronising call).
 */
#ifndef SP_B200_H
#define SP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP_ABI_VERSION 1

/* ---- status codes -------------------------------------------------------- */
enum {
    SP_OK = 0,
    SP_ERR_INVALID = 1,   /* bad argument (null pointer, bad id, wrong field count ...) */
    SP_ERR_CUDA = 2,      /* CUDA runtime error, text in sp_last_error */
    SP_ERR_NO_DEVICE = 3, /* no usable CUDA device: there is NO CPU fallback */
    SP_ERR_STATE = 4,     /* call out of order (e.g. sweep before create_cell_list) */
    SP_ERR_NCCL = 5,
    SP_ERR_NOT_CONVERGED = 6 /* sp_poisson_cg reached maxiter: a SOFT status — P_out holds the last iterate, iters and resid are
                                filled; IterativeSolvers.cg returns the same iterate without an error, so the host bindings hand
                                it back to the caller instead of raising */
};

/* ---- SPH kernel families (src/kernels.jl) -------------------------------- */
enum {
    SP_KERNEL_WENDLAND1 = 1, /* kernels.jl:206-228 */
    SP_KERNEL_WENDLAND2 = 2, /* kernels.jl:108-147 */
    SP_KERNEL_WENDLAND3 = 3, /* kernels.jl:156-204 */
    SP_KERNEL_SPLINE23 = 4,  /* kernels.jl:14-60   */
    SP_KERNEL_SPLINE24 = 5   /* kernels.jl:69-99   */
};
enum { SP_KFUN_W = 0, SP_KFUN_DW = 1, SP_KFUN_RDW = 2, SP_KFUN_DDW = 3 /* wendland3 only */ };

/* ---- registered operators -------------------------------------------------
 * `fields` lists field ids in the role order given; `params` in the order given.
 * x_pq = p.x - q.x, v_pq = p.v - q.v, r = |x_pq| (IEEE, un-fused, core.jl:8-10).
 * All binary operators visit exactly the reference's neighbour set:
 * q in the 9/27 linear-offset cells, !(r > h), q !== p  (core.jl:94-112).
 */
enum {
    /* WCSPH — examples/collapse_dry.jl, collapse3d.jl, cavity_flow.jl */
    SP_OP_BALANCE_OF_MASS = 1,
    /* binary. fields {x, v, rho, Drho}; params {kernel, m, h, two_nu}
       Drho_p += (m*rDw(h,r)) * (dot(x_pq, v_pq) + two_nu*(rho_p - rho_q))
       collapse_dry.jl:112-115, collapse3d.jl:87-90; cavity_flow.jl:92-94 is two_nu = 0 */
    SP_OP_FIND_PRESSURE = 2,
    /* unary. fields {rho, Drho, P}; params {dt, c2, rho0, P0}
       rho += Drho*dt; Drho = 0; P = P0 + c2*(rho - rho0)   (P0 = 0 -> term skipped)
       collapse_dry.jl:123-127, collapse3d.jl:92-96, cavity_flow.jl:96-100 */
    SP_OP_INTERNAL_FORCE = 3,
    /* binary. fields {x, v, P, rho, Dv, type}; params {kernel, m, h, mu, rho0}
       if type_p == 0: ker = m*rDw(h,r);
         Dv_p += (-ker*(P_p/rho_p^2 + P_q/rho_q^2))*x_pq;  Dv_p += (2*ker*mu/rho0^2)*v_pq
       collapse_dry.jl:135-141; collapse3d.jl:98-104 with the correction named in DESIGN.md */
    SP_OP_INTERNAL_FORCE_CAVITY = 4,
    /* binary. fields {x, v, P, rho, Dv, type}; params {m, h, Re, vlid, ylid, lid_type}
       cavity_flow.jl:102-114 (rDwendland2, lid extrapolation, Monaghan viscosity) */
    SP_OP_MOVE = 5,
    /* unary. fields {x, v, Dv, type}; params {dtm}   Dv = 0; if type == 0: x += dtm*v
       collapse_dry.jl:148-153 (dtm = 0.5*dt), collapse3d.jl:106-111 (dtm = dt), cavity_flow.jl:117-122 */
    SP_OP_ACCELERATE = 6,
    /* unary. fields {v, Dv, type}; params {hdt, gx, gy, gz}   if type == 0: v += hdt*(Dv + g)
       collapse_dry.jl:155-159, collapse3d.jl:113-117, cavity_flow.jl:124-128 (g = 0) */

    /* operators of tests/test_collision_2d.jl */
    SP_OP_DENSITY_SUM = 7,
    /* binary, honours self. fields {x, out}; params {kernel, m, h}   out_p += m*w(h,r)
       test_collision_2d.jl:63-69 (find_rho!, find_rho0!) */
    SP_OP_PRESSURE_FROM_RHO = 8,
    /* unary. fields {rho, rho0, P}; params {c2}   P = c2*(rho - rho0)   test_collision_2d.jl:71-73 */
    SP_OP_INTERNAL_FORCE_SYM = 9,
    /* binary. fields {x, P, a}; params {kernel, m, h, rho0}
       a_p += (-(m*rDw)*(P_p/rho0^2 + P_q/rho0^2))*x_pq     test_collision_2d.jl:75-78 */
    SP_OP_FILL = 10,
    /* unary. fields {f}; params {value}   every component of f = value   test_collision_2d.jl:80-86 */
    SP_OP_ADVECT = 11,
    /* unary. fields {x, v}; params {dt}   x += dt*v   test_collision_2d.jl:88-90 */
    SP_OP_KICK = 12,
    /* unary. fields {v, a}; params {hdt}   v += hdt*a   test_collision_2d.jl:92-94 */

    /* ISPH — examples/collapse_dry_implicit.jl */
    SP_OP_ISPH_INITIALIZE = 20,
    /* unary. fields {x, v, div, L, lambda, type}; params {dt, gx, gy, gz}   :118-126 */
    SP_OP_ISPH_VISCOUS_FORCE = 21,
    /* binary. fields {x, v, Dv}; params {kernel, m, h, mu, rho}   Dv_p += (2*m*mu*rDk/rho^2)*v_pq   :128-130 */
    SP_OP_ISPH_DIV_L_LAMBDA = 22,
    /* binary. fields {x, v, div, L, lambda}; params {kernel, m, h, rho, dim}   :147-152 */
    SP_OP_ISPH_PROJECTION_VECTOR = 23,
    /* unary (assemble_vector). fields {div, b}; params {h, dt}   b = -h^2*div/dt   :165-167 */
    SP_OP_ISPH_INTERNAL_FORCE = 25,
    /* binary. fields {x, P, Dv}; params {kernel, m, h, rho}   Dv_p -= (m*rDk*(P_p+P_q)/rho^2)*x_pq   :132-134 */
    SP_OP_ISPH_ACCELERATE = 26,
    /* unary. fields {v, Dv, type}; params {dt}   if type == 0: v += dt*Dv;  Dv = 0   :136-141 */

    /* WCSPH with the density integrated in the pair loop — examples/static_container.jl */
    SP_OP_SC_BALANCE_OF_MASS = 30,
    /* binary. fields {x, v, rho}; params {kernel, m, h, dt}
       rho_p += dt*dot(x_pq, v_pq)*m*rDw(h,r)        static_container.jl:102-104 */
    SP_OP_SC_INTERNAL_FORCE = 31,
    /* binary. fields {x, v, rho, a, type}; params {kernel, m, h, mu, c2, rho0}   with P(rho) = c2*(rho - rho0):
       if type_p == 0: ker = m*rDw(h,r);
         a_p += (-ker*(P(rho_p)/rho_p^2 + P(rho_q)/rho_q^2))*x_pq;  a_p += (ker*2*mu/(rho_p*rho_q))*v_pq
       static_container.jl:106-114 (pressure: :68-70) */
    SP_OP_MOVE_ALL = 32,
    /* unary. fields {x, v, a}; params {dtm}   x += dtm*v (every particle); a = 0   static_container.jl:116-119 */

    /* surface tension by colour-field normals — examples/drop.jl (Wendland quintic 3-D only: uses DDwendland3) */
    SP_OP_FIND_NORMAL = 33,
    /* binary. fields {x, n}; params {kernel, coef, h}   n_p += (coef*rDw(h,r))*x_pq, coef = 2*vol*vol   drop.jl:76-78 */
    SP_OP_NORMALIZE = 34,
    /* unary. fields {n}; params {s0}   s = norm(n); n /= (s + s0)   drop.jl:84-87 */
    SP_OP_INTERNAL_FORCE_TENSION = 35,
    /* binary. fields {x, v, P, n, a}; params {m, h, mu, rho0, beta, s0}   ker = m*rDwendland3(h,r);
         a_p += (-ker*(P_p/rho0^2 + P_q/rho0^2))*x_pq;   a_p += (2*ker*mu/rho0^2)*v_pq;
         a_p -= (2*beta/rho0^2)*(((m*DDwendland3(h,r) - ker)*dot(x_pq, n_pq))*x_pq/(r^2 + s0) + ker*n_pq)
       drop.jl:101-113 */

    /* reversible (symplectic, fixed-point) WCSPH with Lennard-Jones walls — examples/collapse_symplectic.jl,
       examples/Kepler_vortex.jl, examples/utils/FixPA.jl.
       rev_add(a, b) = 2^-30 * (Int64(round(a*2^30)) + Int64(round(b*2^30)))   FixPA.jl:11-42 (round half to even) */
    SP_OP_DENSITY_SUM_FLUID = 40,
    /* binary, honours self. fields {x, out, type}; params {kernel, m, h}
       if type_p == 0 && type_q == 0: out_p += m*w(h,r)
       collapse_symplectic.jl:98-108, Kepler_vortex.jl:139-149 (find_rho!, find_rho0!) */
    SP_OP_INTERNAL_FORCE_LJ = 41,
    /* binary. fields {x, P, rho, a, type}; params {kernel, m, h, rho0, wall_type, dr_wall, E_wall, eps}
       if type_p == 0 && type_q == 0:  a_p += (-(m*rDw(h,r))*(pr_p + pr_q))*x_pq
            with pr = P/rho^2 when rho0 == 0 (collapse_symplectic.jl:114-117), pr = P/rho0^2 otherwise
            (Kepler_vortex.jl:155-158)
       else if type_p == 0 && type_q == wall_type && r < dr_wall:
            s = dr_wall/(r + eps);  a_p += (-E_wall/(r + eps)^2*(s^2 - s^4))*x_pq
       collapse_symplectic.jl:114-123, Kepler_vortex.jl:155-164 */
    SP_OP_MOVE_REV = 42,
    /* unary. fields {x, v, type}; params {dt}   if type == 0: x = rev_add(x, dt*v)
       collapse_symplectic.jl:134-138, Kepler_vortex.jl:174-178 */
    SP_OP_ACCELERATE_REV = 43,
    /* unary. fields {v, a, type}; params {hdt, gx, gy, gz}   if type == 0: v = rev_add(v, hdt*(a + g))
       collapse_symplectic.jl:140-144 */
    SP_OP_ACCELERATE_REV_CENTRAL = 44,
    /* unary. fields {x, v, a, type}; params {hdt, GM}
       if type == 0: v = rev_add(v, hdt*rev_add(a, (-GM/norm(x)^3)*x))   Kepler_vortex.jl:180-184 */
    SP_OP_LJ_POTENTIAL = 45,
    /* binary (the per-particle sum(sys, LJ_potential, p), core.jl:271-291). fields {x, out, type};
       params {h, coef, wall_type, dr_wall, eps}
       if type_q == wall_type && type_p == 0 && r < dr_wall: s = dr_wall/(r + eps);
            out_p += coef*(0.5*s^2 - 0.25*s^4 - 0.25),  coef = m*E_wall
       collapse_symplectic.jl:146-153, Kepler_vortex.jl:186-193 */

    /* channel flow past a cylinder with an inflow buffer and per-particle mass — examples/cylinder.jl */
    SP_OP_CYL_BALANCE_OF_MASS = 50,
    /* binary. fields {x, v, rho, Drho, m, type}; params {kernel, h, two_nu}
       Drho_p += (m_q*rDw(h,r))*dot(x_pq, v_pq);  if type_p == 0 && type_q == 0: Drho_p += two_nu/rho_p*(rho_p - rho_q)
       cylinder.jl:102-108 */
    SP_OP_CYL_FIND_PRESSURE = 51,
    /* unary. fields {x, rho, Drho, P}; params {dt, c2, rho0, x1_min}
       if x[1] >= x1_min: rho += Drho*dt;   Drho = 0;  P = c2*(rho - rho0)       cylinder.jl:110-116 */
    SP_OP_CYL_INTERNAL_FORCE = 52,
    /* binary. fields {x, v, P, rho, a, m}; params {kernel, h, mu, eps2}   ker = m_q*rDw(h,r);
       a_p += (-ker*(P_p/rho_p^2 + P_q/rho_q^2))*x_pq;
       a_p += (8*ker*mu/(rho_p*rho_q)*dot(v_pq, x_pq)/(r*r + eps2))*x_pq,  eps2 = 0.01*h*h      cylinder.jl:118-123 */
    SP_OP_MOVE_TYPES = 53,
    /* unary. fields {x, v, a, type}; params {dt, type_a, type_b}
       a = 0;  if type == type_a || type == type_b: x += dt*v                     cylinder.jl:125-130 */
    SP_OP_CYL_ACCELERATE = 54,
    /* unary. fields {x, v, a, type}; params {hdt, cyl1, coef}   if type == 0:
       f = (cyl1 - x[1], -x[2], 0);  v += hdt*(a + coef*f/((cyl1 - x[1])^2 + x[2]^2))   cylinder.jl:132-143 */
    SP_OP_SET_INFLOW_SPEED = 55,
    /* unary. fields {x, v, type}; params {inflow_type, s, U_max, chan_w}
       if type == inflow_type: v = (s*U_max*(1 - (2*x[2]/chan_w)^2))*VECX,  s = min(1, t/t_acc) from the host
       cylinder.jl:91-97 */

    /* elastic solid with tensor-valued particle fields — examples/rod.jl.  A, H, B are 9-component fields holding a
       RealMatrix in Julia's column-major order (component c = (i-1) + 3*(j-1) for M[i,j]); the script's own 2-D
       outer/det/inv/trans/dev (rod.jl:44-85) only populate the in-plane block M[1:2,1:2], which is what is computed. */
    SP_OP_ROD_FIND_A = 60,
    /* binary. fields {x, X, A, H}; params {kernel, h}   ker = w(h,r):
       A_p += -ker*outer(X_pq, x_pq);  H_p += -ker*outer(x_pq, x_pq)              rod.jl:128-134 */
    SP_OP_ROD_FIND_B = 61,
    /* unary. fields {A, H, B}; params {m, c_l, c_s}   Hi = inv(H); A = A*Hi; At = trans(A); G = At*A;
       P = c_l^2*(det(A) - 1);  B = m*(P*inv(At) + c_s^2*A*dev(G))*Hi              rod.jl:136-143 */
    SP_OP_ROD_FIND_F = 62,
    /* binary. fields {x, v, X, A, B, f}; params {kernel, h, two_m_vol, nu}        rod.jl:145-160
       f_p += -ker*(A_p'*(B_p*x_pq)) - ker*(A_q'*(B_q*x_pq)) + eta-correction terms + (two_m_vol*rDker*nu)*v_pq */
    SP_OP_ROD_PULL = 63,
    /* unary. fields {X, f}; params {X1_min, fy}   if X[1] > X1_min: f += (0, fy, 0)      rod.jl:162-166 */
    SP_OP_ROD_UPDATE_V = 64,
    /* unary. fields {v, f, X}; params {hdt, m, X1_clamp}   v += hdt*f/m;  if X[1] < X1_clamp: v = 0   rod.jl:168-174 */
    SP_OP_ROD_UPDATE_X = 65,
    /* unary. fields {x, v, A, H, f, e}; params {dt}   x += dt*v;  H = A = 0; f = 0; e = 0      rod.jl:176-183 */
    SP_OP_ROD_FIND_E = 66,
    /* binary. fields {x, X, A, e}; params {h}   eta = inv(A_p)*X_pq - x_pq;  e_p += dot(eta, eta)   rod.jl:185-188 */

    /* SHTC fluid (full 3x3 distortion field A, stress tensor) — examples/SHTC/ldc.jl.  Parity on B200:
       tests/test_shtc_gpu.py; the oracle side is pinned in tests/test_shtc_cpu.py. */
    SP_OP_SHTC_FIND_STRESS = 70,
    /* unary. fields {A, rho, stress}; params {c_l, c_s, rho_ref}   G = A'*A;
       stress = c_l^2*(rho - rho_ref)*I + c_s^2*rho*G*dev(G),  rho_ref = rho0/(1 + acf)        ldc.jl:118-121 */
    SP_OP_SHTC_UPDATE_V = 71,
    /* binary. fields {x, v, rho, stress, type}; params {kernel, h, dtm}   if type_p == 0:
       v_p += ((-dtm*rDw(h,r))*(stress_p/rho_p^2 + stress_q/rho_q^2))*x_pq,  dtm = dt*m             ldc.jl:123-127 */
    SP_OP_SHTC_UPDATE_RHO = 72,
    /* binary. fields {x, v, rho, type}; params {kernel, h, dtm}   if type_p == 0:
       rho_p += (dtm*rDw(h,r))*dot(x_pq, v_pq)                                                     ldc.jl:90-94 */
    SP_OP_SHTC_CONVECT_A = 73,
    /* binary, ORDER-DEPENDENT (each pair uses the A_p the previous pairs left): always swept in the reference's
       visiting order.  fields {x, v, rho, A, type}; params {kernel, h, dtm, skip_type}   if type_p != skip_type:
       A_p += ((dtm/rho_p*rDw(h,r))*A_p)*(v_pq*x_pq')                                              ldc.jl:96-100 */
    SP_OP_SHTC_RELAX_A = 74,
    /* unary. fields {A}; params {dt, tau}   one RK4 step of dA/dt = -3/tau*A*dev(A'*A)             ldc.jl:102-116 */
    SP_OP_SHTC_MOVE = 75,
    /* unary. fields {x, v, type}; params {dt}   if type == 0: x += v*dt                            ldc.jl:129-133 */

    /* SHTC solid in 2-D (vibrating beryllium plate) — examples/SHTC/beryllium.jl.  T, L, A are RealMatrix fields; the
       script's own outer/det/inv/dev (:79-103) are the 2-D ones (inv sets [3,3] = 1).  w_h / rDw_h are the script's
       "structural" kernels wendland2h / rDwendland2h (:44-52, strict x < 1).  Parity on B200:
       tests/test_shtc_gpu.py; the oracle side is pinned in tests/test_shtc_cpu.py.  update_x! is SP_OP_ADVECT. */
    SP_OP_BE_FIND_L = 80,
    /* binary. fields {x, v, m, T, L}; params {kernel, h, rho0}   ker = m_q/rho0*rDw(h,r):
       T_p += ker*outer(x_pq, x_pq);  L_p += ker*outer(v_pq, x_pq)                              beryllium.jl:140-146 */
    SP_OP_BE_UPDATE_A = 81,
    /* unary. fields {A, T, L}; params {hdt}   L = L*inv(T);  A = A*(I - hdt*L)*inv(I + hdt*L)   beryllium.jl:148-151 */
    SP_OP_BE_FIND_J = 82,
    /* binary, honours self. fields {x, m, T, J, K}; params {kernel, h, rho0}   T_p += (m_q/rho0*rDw)*outer(x_pq, x_pq);
       J_p += m_q/rho0*w(h,r);  K_p += m_q/rho0*w_h(h,r)        beryllium.jl:153-158; taco.jl:141-146 (find_rho!, self) */
    SP_OP_BE_FIND_T = 83,
    /* unary. fields {A, T, P, J}; params {rho0, c_0, c_s}   G = A'*A;
       P = 0.5*rho0*c_0^2*((1 - 1/J)/J^2 + log(J)/J);  T = P/rho0*I - c_s^2*G*dev(G)*inv(T)       beryllium.jl:160-164 */
    SP_OP_BE_FIND_F = 84,
    /* binary. fields {x, m, T, K, f}; params {kernel, h, rho0, c_p}   ker = m_q/rho0*rDw, kerh = m_q/rho0*rDw_h:
       f_p += -m_p*ker*(T_p*x_pq) - m_p*ker*(T_q*x_pq) - (m_p*kerh*c_p^2*(K_p + K_q))*x_pq      beryllium.jl:166-175 */
    SP_OP_BE_RESET = 85,
    /* unary. fields {f, L, T, J, K, J0, K0}   f = 0; L = 0; T = 0; J = J0; K = K0                 beryllium.jl:177-184 */
    SP_OP_BE_UPDATE_V = 86,
    /* unary. fields {v, f, m}; params {hdt}   v += hdt*f/m                                        beryllium.jl:132-134 */

    /* SHTC solid in 3-D (twisting column) — examples/SHTC/twist3d.jl: the beryllium operators with full 3x3 matrices,
       StaticArrays' general inverse and the 3-D structural kernels wendland3h / rDwendland3h (:43-51).  Parity on B200:
       tests/test_shtc_gpu.py; the oracle side is pinned in tests/test_shtc_cpu.py.
       reset! is SP_OP_BE_RESET, update_x! is SP_OP_ADVECT. */
    SP_OP_TW_FIND_L = 90,
    /* binary. fields {x, v, m, T, L}; params {kernel, h, rho0}                                  twist3d.jl:135-141 */
    SP_OP_TW_UPDATE_A = 91,
    /* unary. fields {A, T, L}; params {hdt}   L = L*inv(T);  A = A*(I - hdt*L)*inv(I + hdt*L)   twist3d.jl:143-146 */
    SP_OP_TW_FIND_J = 92,
    /* binary. fields {x, m, T, J, K}; params {kernel, h, rho0}                                  twist3d.jl:148-153 */
    SP_OP_TW_FIND_T = 93,
    /* unary. fields {A, T, P, J}; params {rho0, c_0, c_s}   F = inv(A); B = F*F'; detF = 1/J;
       P = -rho0*c_0^2*detF^2*(detF - 1);  T = -P/rho0*I - c_s^2*(B - I)*inv(T)                    twist3d.jl:155-161 */
    SP_OP_TW_FIND_F = 94,
    /* binary. fields {x, m, T, K, f}; params {kernel, h, rho0, c_p}
       f_p += m_p*ker*(T_p*x_pq) + m_p*ker*(T_q*x_pq) - (m_p*kerh*c_p^2*(K_p + K_q))*x_pq          twist3d.jl:163-172 */
    SP_OP_TW_UPDATE_V = 95,
    /* unary. fields {x, v, f, m}; params {hdt}   if x[3] > 0: v += hdt*f/m                        twist3d.jl:125-129 */

    /* SHTC fluid between rotating cylinders (Taylor-Couette) — examples/SHTC/taco.jl.  find_L!, update_A!, reset! and
       find_rho! (with self) are SP_OP_BE_FIND_L / BE_UPDATE_A / BE_RESET / BE_FIND_J with rho0 = 1 (the same arithmetic:
       m/1.0 is m), relax_A! is SP_OP_SHTC_RELAX_A.  Parity on B200: tests/test_shtc_gpu.py. */
    SP_OP_TA_FIND_T = 100,
    /* unary. fields {A, T, P, rho}; params {rho0, c_0, c_s}   G = A'*A;  P = c_0^2*(rho - rho0)*rho0/rho;
       T = -P/rho^2*I + c_s^2*G*dev(G)*subinv(T)                                                  taco.jl:148-152 */
    SP_OP_TA_FIND_F = 101,
    /* binary. fields {x, m, T, lambda, f}; params {kernel, h, cpr2}   ker = m_q*rDw, kerh = m_q*rDw_h:
       f_p += (m_p*ker*(T_p + T_q))*x_pq - (m_p*kerh*cpr2*(lambda_p + lambda_q))*x_pq,  cpr2 = (c_p/rho0)^2   taco.jl:154-162 */
    SP_OP_TA_UPDATE_V = 102,
    /* unary. fields {x, v, f, m, type}; params {hdt, R1, R2, omega}   if type == 0: v += hdt*f/m
       else v = R2/r*(r/R1 - R1/r)/(R2/R1 - R1/R2)*(-omega*x[2], omega*x[1], 0), r = norm(x)       taco.jl:108-114, 39-42 */
    SP_OP_TA_UPDATE_X = 103
    /* unary. fields {x, v, x0, type}; params {hdt, cos_wt, sin_wt, outer_type}   if type == 0: x += hdt*v;
       if type == outer_type: x = (x0[1]*cos_wt - x0[2]*sin_wt, x0[1]*sin_wt + x0[2]*cos_wt, 0)    taco.jl:116-126 */
};

/* sp_apply flags */
enum {
    SP_FLAG_SELF = 1,        /* apply!(...; self=true): add the (p,p,0.0) term after the sweep (core.jl:155-157) */
    SP_FLAG_STRICT_ORDER = 2, /* accumulate in the reference's order: key_diff order x descending index */
    SP_FLAG_TILE_KERNEL = 4,  /* alternative kernel: shared-memory tile (TMA bulk staging + FP32 pre-filter), no lists */
    SP_FLAG_UNFUSED_BUILD = 8 /* build the neighbour lists with their own kernel even when the operator has the fused
                                 build + first-replay kernel (A/B timing, parity of the two-kernel path) */
};

/* reductions (diagnostic loops of the examples) */
enum {
    SP_RED_ENERGY_WCSPH = 1,
    /* fields {x, v, rho}; params {m, c, rho0, gx, gy, gz}; out[1]
       sum of 0.5 m v.v - m g.x + m c^2 (log|rho/rho0| + rho0/rho - 1)   collapse_dry.jl:166-171 */
    SP_RED_FRONT = 2,
    /* fields {x, type}; params {width, height, h, xmax}; out[2] = {X, H}   collapse_dry.jl:173-187 */
    SP_RED_ENERGY_COLLISION = 3,
    /* fields {v, rho, rho0}; params {m, c, rho0}; out[1]   test_collision_2d.jl:96-100 */
    SP_RED_SUM = 4,
    /* fields {f}; out[ncomp] plain sum of every component */
    SP_RED_ENERGY_ISPH = 5,
    /* fields {x, v}; params {m, gx, gy, gz}; out[1]   collapse_dry_implicit.jl:173-177 */
    SP_RED_FORCE_ON_TYPE = 6,
    /* fields {a, m, type}; params {type_sel}; out[3] = sum of m*a over the particles with type == type_sel
       cylinder.jl:158-159 (calculate_force over the obstacle particles) */
    SP_RED_ENERGY_ROD = 7,
    /* fields {v, A}; params {m, c_s, c_l}; out[1]   sum of 0.5 m v.v + 0.25 m c_s^2 |dev(A'A)|_F^2
       + m c_l^2 (d - 1 - log d), d = |det A|                                       rod.jl:190-199 */
    SP_RED_MAX_SPEED = 8
    /* fields {v}; out[1] = max over the particles of norm(v) (0 for an empty system).  The ingredient of an adaptive
       CFL time step dt = cfl*h/(c + max|v|); on a slab system the maximum is all-reduced over the ranks (NCCL max),
       so every rank gets the same dt.  No counterpart in the reference: its examples use fixed time steps. */
};

/* point sums:  out[k] = sum_q func(q, |x_k - q.x|)  over !(r > h), no self exclusion (core.jl:240-260) */
enum {
    SP_SUM_MASS_W = 1,  /* fields {x, type}; params {kernel, m, h, type_sel}: [type==type_sel]*m*w(h,r)       cavity_flow.jl:169,173 */
    SP_SUM_MASS_F_W = 2 /* fields {x, type, f}; params {kernel, m, h, type_sel, comp}: ... *f[comp]*w(h,r)   cavity_flow.jl:170,174 */
};

/* host array layout for upload/download of a field with ncomp components */
enum {
    SP_LAYOUT_AOS = 0, /* host[i*ncomp + c]  (a Julia Vector{SVector{3,Float64}} reinterpreted) */
    SP_LAYOUT_SOA = 1  /* host[c*n + i] */
};

typedef struct sp_system sp_system;

/* ---- library ------------------------------------------------------------- */
int32_t sp_version(void);
/* Text of the last error on this handle (or of the last failed sp_create when sys == NULL). */
const char* sp_last_error(const sp_system* sys);
int32_t sp_device_count(int32_t* count);

/* ---- particle system (src/structs.jl:57-91) ------------------------------
 * lo/hi = corners of boundarybox(domain); h = neighbour radius.  Computes key_phase,
 * key_lim, key_max, key_diff exactly as the constructor does.  Field 0 "x" (3 comps)
 * always exists. */
int32_t sp_create(sp_system** out, const double lo[3], const double hi[3], double h, int32_t device);
int32_t sp_destroy(sp_system* sys);
int32_t sp_key_params(const sp_system* sys, int64_t key_phase[3], int64_t key_lim[3], int64_t* key_max,
                      int32_t* n_key_diff, int64_t key_diff[27]);

/* ---- fields (struct-of-arrays replacement of the particle struct) -------- */
int32_t sp_add_field(sp_system* sys, const char* name, int32_t ncomp, int32_t* fid);
int32_t sp_find_field(const sp_system* sys, const char* name, int32_t* fid);
/* Set the particle count.  Growing appends zero-initialised particles at the end of the
 * reference order (push!, src/grids.jl:256); shrinking drops the tail. */
int32_t sp_resize(sp_system* sys, int64_t n);
int32_t sp_num_particles(sp_system* sys, int64_t* n);
/* Host data is always in REFERENCE ORDER (index i = sys.particles[i+1] of the reference),
 * whatever order the device keeps internally.  n must equal the particle count. */
int32_t sp_upload(sp_system* sys, int32_t fid, const double* host, int64_t n, int32_t layout);
int32_t sp_download(sp_system* sys, int32_t fid, double* host, int64_t n, int32_t layout);
int32_t sp_synchronize(sp_system* sys);

/* ---- generate_particles! on the device (src/grids.jl:52-144, 253-258; shapes src/geometry.jl:15-258) ----------
 * Every lattice point of the index box irange = {i0,i1,j0,j1,k0,k1} (inclusive; what floor/ceil of the shape's
 * boundarybox over the lattice constant give, grids.jl:53-56) is tested against the shape; the points inside are
 * appended to the system in the reference's generation order with the same Float64 coordinates.  The shape is a
 * postfix program (children before parents, root last).  fill_fields/fill_values set constant scalar or vector
 * fields of the new particles (what the example's particle constructor does, e.g. type, rho0); the rest is zero. */
typedef struct {
    int32_t kind;  /* SP_SHAPE_* */
    int32_t a, b;  /* child node indices; HALFSPACE: a = axis (0,1,2), b = 0 '<', 1 '<=', 2 '>', 3 '>=' */
    double p[8];   /* BOX lo[3],hi[3] | CIRCLE cx,cy,r*r | BALL cx,cy,cz,r*r | HALFSPACE bound */
} sp_shape_node;
enum {
    SP_SHAPE_BOX = 1,            /* Box / Rectangle, closed            geometry.jl:15-43 */
    SP_SHAPE_CIRCLE = 2,         /* geometry.jl:49-68 */
    SP_SHAPE_BALL = 3,           /* geometry.jl:245-258 */
    SP_SHAPE_UNION = 4,          /* geometry.jl:108-127 */
    SP_SHAPE_INTERSECTION = 5,   /* geometry.jl:134-153 */
    SP_SHAPE_DIFFERENCE = 6,     /* geometry.jl:160-171 */
    SP_SHAPE_HALFSPACE = 7,      /* the Specification predicates of the examples (x[2] < h, ...)  geometry.jl:178-189 */
    SP_SHAPE_BOUNDARY_LAYER = 8  /* geometry.jl:198-234: not in a, but x + dx in a for a lattice offset |dx| <= width */
};
enum { SP_GRID_SQUARE = 1, SP_GRID_HEXAGONAL = 2, SP_GRID_CUBIC = 3 }; /* grids.jl:48-91, 124-144 */
int32_t sp_generate_particles(sp_system* sys, int32_t grid, double dr, const sp_shape_node* nodes, int32_t n_nodes,
                              const double* offsets /* n_off x 3 */, int32_t n_off, const int64_t irange[6],
                              const int32_t* fill_fields, const double* fill_values, int32_t n_fill, int64_t* n_added);

/* Inflow buffer, add_new_particles! of examples/cylinder.jl:145-156: every particle (in reference order) with
 * type == from_type and x[1] >= x1_min becomes to_type, and a new particle of type from_type is appended at
 * x - shift*VECX; the appended particles keep the order of their sources.  New particles are what the script's
 * constructor makes: every field zero except the `n_fill` constant fields (e.g. rho = rho0, m = m0) and `type`. */
int32_t sp_respawn(sp_system* sys, int32_t type_field, double from_type, double to_type, double x1_min, double shift,
                   const int32_t* fill_fields, const double* fill_values, int32_t n_fill, int64_t* n_added);

/* ---- the hot path --------------------------------------------------------- */
int32_t sp_create_cell_list(sp_system* sys);
int32_t sp_apply(sp_system* sys, int32_t op, const int32_t* fields, int32_t nfields, const double* params,
                 int32_t nparams, int32_t flags);
int32_t sp_sum_at_points(sp_system* sys, int32_t sum_op, const int32_t* fields, int32_t nfields,
                         const double* params, int32_t nparams, const double* xyz /* m x 3 AoS */, int64_t m,
                         double* out /* m */);
int32_t sp_reduce(sp_system* sys, int32_t red, const int32_t* fields, int32_t nfields, const double* params,
                  int32_t nparams, double* out);

/* ---- ISPH pressure-Poisson operator, matrix-free ---------------------------
 * (A p)_i = A_ii p_i + sum_{j != i, r <= h} (2 h^2 m/rho) rDk(h,r_ij) p_j,
 * A_ii = h^2 L_i + [type_i == 0] C_free max(lambda_i, 0)
 * = the matrix assemble_matrix(sys, projection_matrix) builds (collapse_dry_implicit.jl:154-163).
 * fields {x, L, lambda, type, p_in, y_out}; params {kernel, m, h, rho, C_free}. */
int32_t sp_poisson_apply(sp_system* sys, const int32_t* fields, int32_t nfields, const double* params,
                         int32_t nparams);
/* Unpreconditioned CG from x0 = 0 on A P = b, stopping at |r| <= max(reltol*|b|, abstol) or maxiter
 * (IterativeSolvers.cg defaults: reltol = sqrt(eps), abstol = 0, maxiter = N).
 * fields {x, L, lambda, type, b, P}; the result is written to P.  maxiter <= 0 -> N. */
int32_t sp_poisson_cg(sp_system* sys, const int32_t* fields, int32_t nfields, const double* params,
                      int32_t nparams, double reltol, double abstol, int64_t maxiter, int64_t* iters,
                      double* resid);

/* assemble_matrix(sys, projection_matrix) (src/core.jl:196-225 with collapse_dry_implicit.jl:154-163) for a host that
 * wants the matrix itself (a direct solver, a preconditioner): COO triplets, 1-based reference indices, one triplet
 * per particle for the diagonal and one per (particle, neighbour within h) — the triplets the reference hands to
 * sparse(I, J, V, N, N), which sums duplicates (the narrow-domain double visit).  Two-call protocol: with
 * I == J == V == NULL only *nnz is written; otherwise `cap` entries are available and *nnz are filled.
 * fields {x, L, lambda, type}; params {kernel, m, h, rho, C_free}.  Single-GPU systems only. */
int32_t sp_assemble_matrix(sp_system* sys, const int32_t* fields, int32_t nfields, const double* params, int32_t nparams,
                           int64_t* I, int64_t* J, double* V, int64_t cap, int64_t* nnz);

/* ---- fused step programs (amortise launch latency; same arithmetic) ------ */
enum {
    SP_PROGRAM_WCSPH_3D = 1, /* examples/collapse3d.jl:136-150 — move, cell list, balance_of_mass,
                                find_pressure, internal_force, accelerate, accelerate */
    SP_PROGRAM_WCSPH_2D = 2  /* examples/collapse_dry.jl:203-211 */
};
/* fields {x, v, Dv, rho, Drho, P, type}; params {kernel, m, h, two_nu, dt, c2, rho0, mu, gx, gy, gz} */
int32_t sp_run_program(sp_system* sys, int32_t program, const int32_t* fields, int32_t nfields,
                       const double* params, int32_t nparams, int64_t nsteps);

/* ---- step graphs: any host loop body as ONE launch -------------------------------------------------------------
 * No counterpart in the reference; the device analogue of "the loop body is cheap to call".  The cell-list build keeps
 * its counts on the device, so a time step of ANY example is a fixed sequence of kernel launches.  Between
 * sp_graph_begin and sp_graph_end the asynchronous entry points (sp_apply, sp_create_cell_list, sp_poisson_apply,
 * sp_run_program without its own graphs, ...) are recorded into a CUDA graph instead of being executed; sp_graph_end
 * executes the recorded body ONCE — so begin/end behaves like the calls it encloses — and returns a handle;
 * sp_graph_launch(id, times) replays it.  A 10 k-particle 2-D step costs ~25 launches of ~3 us each when issued call
 * by call and one graph launch when replayed.
 * Rules: (1) run the body once the ordinary way first (lazy allocations happen then); (2) the body must leave the
 * library's ping-pong buffers where it found them, which means an EVEN number of cell-list builds — record two time
 * steps if a step has one; otherwise sp_graph_end executes the body, keeps no graph and returns SP_ERR_STATE;
 * (3) calls that hand data or counts to the host (sp_download, sp_reduce, sp_num_particles, sp_poisson_cg, ...) are
 * refused inside a recording (SP_ERR_STATE); (4) a graph is replayed only while everything its launches depend on is
 * unchanged (field storage, slot bound, list capacity): sp_graph_launch returns SP_ERR_STATE otherwise and the host
 * records again.  Particles leaving the domain inside a replay are handled on the device as in the ordinary path. */
int32_t sp_graph_begin(sp_system* sys);
int32_t sp_graph_end(sp_system* sys, int32_t* graph_id);
int32_t sp_graph_launch(sp_system* sys, int32_t graph_id, int64_t times);
int32_t sp_graph_destroy(sp_system* sys, int32_t graph_id);

/* ---- kernel functions on the device (tests/test_kernels.jl parity) ------- */
int32_t sp_kernel_eval(int32_t kernel, int32_t kfun, double h, const double* r, double* out, int64_t n,
                       int32_t device);

/* ---- parity / debug views (all in reference order, 1-based like the reference) */
/* keys[i] = find_key(sys, particles[i].x) as stored by the last sp_create_cell_list */
int32_t sp_get_cell_keys(sp_system* sys, int64_t* keys, int64_t n);
/* CSR view of cell_list: cell k (1-based) holds members[offsets[k-1] .. offsets[k]) — 1-based particle
 * indices in DESCENDING order, exactly the non-zero prefix of cell_list[k].entries (core.jl:26-41). */
int32_t sp_get_cell_list(sp_system* sys, int64_t* offsets /* key_max+1 */, int64_t* members /* n */);
/* Neighbour lists as visited by apply_binary!: for particle i (0-based slot in the arrays),
 * ids[offsets[i] .. offsets[i+1]) = 1-based indices q in the reference's visiting order.
 * Call with ids == NULL to get only offsets (offsets[n] = total). */
int32_t sp_get_neighbour_lists(sp_system* sys, int64_t* offsets /* n+1 */, int64_t* ids, int64_t ids_cap);
/* Optional: build the cached neighbour lists of the current positions now (the first binary operator after the
 * positions change does it implicitly).  The op-independent part of apply_binary! (core.jl:94-110: key, candidate
 * cells, dist, r > h, identity) runs once here; every following sp_apply of a binary operator replays the lists
 * until a position write, re-sort or resize invalidates them. */
int32_t sp_build_neighbour_lists(sp_system* sys);
/* Entries per target the cached lists currently hold (64 at first).  A build that meets a target with more neighbours
 * sweeps that target with the exact candidate scan, reports its longest list, and the next build allocates
 * 1.25 x that (multiple of 32, at most 512): dense kernels such as h = 3 dr in 3-D (~113 neighbours) reach the list
 * path after the first step. */
int32_t sp_neighbour_list_capacity(sp_system* sys, int32_t* capk);
/* The neighbour lists the default pair sweeps actually replay (the per-position-version cache built by the first
 * sweep after positions change; built here if needed).  Same id SET per particle as sp_get_neighbour_lists; the
 * order is the default sweep's visiting order (stencil rows dk,dj outer, slots ascending), not the reference's. */
int32_t sp_get_sweep_neighbour_lists(sp_system* sys, int64_t* offsets /* n+1 */, int64_t* ids, int64_t ids_cap);
/* Number of particles removed by all sp_create_cell_list calls so far. */
int32_t sp_num_removed(sp_system* sys, int64_t* n_removed);
/* Device timing of the last call in ms (CUDA events on the handle's stream). */
int32_t sp_last_call_ms(sp_system* sys, float* ms);
/* Time a region of calls on the handle's stream with CUDA events: start records, stop records,
 * synchronises and returns the elapsed device time in ms. */
int32_t sp_timer_start(sp_system* sys);
int32_t sp_timer_stop(sp_system* sys, float* ms);
/* Count of kernel launches issued by this handle since creation. */
int32_t sp_launch_count(sp_system* sys, int64_t* launches);

/* ---- slab decomposition over the GPUs of one node (no counterpart in the reference: SURVEY 8(e)) ----
 * One process per GPU.  Rank 0 obtains a 128-byte NCCL unique id (sp_slab_unique_id), the host passes it to
 * the other ranks out of band, and every rank calls sp_slab_init on a system created with the SAME global box
 * and h, BEFORE adding particles.  The global cell grid is cut along the slowest key axis (z in 3-D, y in 2-D)
 * into `nranks` slabs of whole cell layers; each rank must be given the particles whose cell layer lies in
 * its slab (sp_slab_range).  periodic != 0 wraps the slab axis (only that axis) with period key_lim[axis]*h;
 * the global box must then start and end on cell faces along that axis. */
int32_t sp_slab_unique_id(uint8_t id[128]);
int32_t sp_slab_init(sp_system* sys, const uint8_t id[128], int32_t rank, int32_t nranks, int32_t periodic);
/* The same with the caller's cut planes instead of equal layer counts: rank r owns the cell layers [cuts[r], cuts[r+1])
 * (0-based from the global key_phase of the slab axis; cuts[0] = 0, cuts[nranks] = key_lim[axis]; at least three layers
 * per rank).  A host that knows how many particles each layer holds balances the particle counts with it. */
int32_t sp_slab_init_cuts(sp_system* sys, const uint8_t id[128], int32_t rank, int32_t nranks, int32_t periodic,
                          const int64_t* cuts);
/* Owned cell layers [cell_lo, cell_hi) (0-based from the global key_phase) and their coordinate interval. */
int32_t sp_slab_range(sp_system* sys, int64_t* cell_lo, int64_t* cell_hi, double* coord_lo, double* coord_hi,
                      int32_t* axis);
/* create_cell_list! for a slab system: ONE exchange round — every owned particle whose current position lies in the
 * rank's first / last two owned cell layers or beyond goes to the lower / upper neighbour (ghost copies and migrants
 * alike, every non-zero field); ownership afterwards follows from the position; then the local cell-list build on the
 * owned layers plus two ghost layers per side.  The ranks of a node write these messages straight into each other's
 * receive buffers (CUDA IPC over NVLink; SP_SLAB_P2P=0 or an unmappable link: NCCL send/recv); no host synchronisation
 * in the steady state, all ranks must issue the same calls (a rank whose neighbour stops answering reports SP_ERR_STATE
 * from a later rebuild, after SP_SLAB_TIMEOUT_S seconds — default 30 — of waiting on the device).  Afterwards the particle count includes the ghosts; field
 * "_ghost" is 0 for owned particles, 1 / 2 for ghosts below / above the owned layers.  Particle order on a slab system
 * is the device order (no reference numbering exists across ranks): identify particles by a field of your own (e.g. a
 * global id).  With two ghost layers the inner one integrates its own density in the WCSPH loops, so those need no
 * further exchange inside a step.  At least three cell layers per rank. */
int32_t sp_slab_create_cell_list(sp_system* sys);
/* Refresh the ghost copies of the listed fields from their owners — for solvers whose vectors change between rebuilds
 * (the CG search direction; sp_poisson_cg does it itself).  A copy of slot ranges: boundary layers and ghost layers hold
 * the same particles in the same order. */
int32_t sp_slab_halo_refresh(sp_system* sys, const int32_t* fields, int32_t nfields);
/* Owned (non-ghost) particle count on this rank after the last sp_slab_create_cell_list. */
int32_t sp_slab_num_owned(sp_system* sys, int64_t* n_owned);
/* Sum / max all-reduce of host doubles over the slab communicator (diagnostics, time-step control). */
int32_t sp_slab_allreduce(sp_system* sys, double* inout, int32_t count, int32_t is_max);

#ifdef __cplusplus
}
#endif
#endif /* SP_B200_H */
