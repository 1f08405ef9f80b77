// sp_program.cu — fused step programs: the time loops of the examples issued from inside the library,
// so a Julia/Python host pays one FFI crossing per batch of steps instead of 7-9 per step.
// Same kernels, same order, same arithmetic as the per-call path.
#include "sp_internal.cuh"

int sp_apply_impl(sp_system* s, int32_t op, const int32_t* F, int32_t nf, const double* Pm, int32_t np, int32_t flags);
int sp_build_cells(sp_system* s);

// fields {x, v, Dv, rho, Drho, P, type}; params {kernel, m, h, two_nu, dt, c2, rho0, mu, gx, gy, gz}
extern "C" int32_t sp_run_program(sp_system* s, int32_t program, const int32_t* F, int32_t nf, const double* P,
                                  int32_t np, int64_t nsteps) {
    if (!s) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    const int nc[] = {3, 3, 3, 1, 1, 1, 1};
    int rc = sp_check_fields(s, F, nf, nc, 7);
    if (rc) return rc;
    if (np != 11 || !P) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters for this program");
    if (s->slab) return sp_fail(s, SP_ERR_STATE, "step programs on slab systems are driven by the host");
    const int32_t x = F[0], v = F[1], Dv = F[2], rho = F[3], Drho = F[4], Pr = F[5], ty = F[6];
    const double kernel = P[0], m = P[1], h = P[2], two_nu = P[3], dt = P[4], c2 = P[5], rho0 = P[6], mu = P[7];
    const int32_t f_bom[4] = {x, v, rho, Drho}, f_fp[3] = {rho, Drho, Pr}, f_if[6] = {x, v, Pr, rho, Dv, ty},
                  f_mv[4] = {x, v, Dv, ty}, f_ac[3] = {v, Dv, ty};
    const double p_bom[4] = {kernel, m, h, two_nu}, p_fp[4] = {dt, c2, rho0, 0.0}, p_if[5] = {kernel, m, h, mu, rho0},
                 p_ac[4] = {0.5 * dt, P[8], P[9], P[10]};
    if ((rc = sp_time_begin(s))) return rc;
#define STEP(call) \
    if ((rc = (call))) return rc;
    for (int64_t k = 0; k < nsteps; k++) {
        if (program == SP_PROGRAM_WCSPH_3D) {  // examples/collapse3d.jl:136-150
            const double p_mv[1] = {dt};
            STEP(sp_apply_impl(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0));
            STEP(sp_build_cells(s));
            STEP(sp_apply_impl(s, SP_OP_BALANCE_OF_MASS, f_bom, 4, p_bom, 4, 0));
            STEP(sp_apply_impl(s, SP_OP_FIND_PRESSURE, f_fp, 3, p_fp, 4, 0));
            STEP(sp_apply_impl(s, SP_OP_INTERNAL_FORCE, f_if, 6, p_if, 5, 0));
            STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
            STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
        } else if (program == SP_PROGRAM_WCSPH_2D) {  // examples/collapse_dry.jl:203-211
            const double p_mv[1] = {0.5 * dt};
            STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
            STEP(sp_apply_impl(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0));
            STEP(sp_build_cells(s));
            STEP(sp_apply_impl(s, SP_OP_BALANCE_OF_MASS, f_bom, 4, p_bom, 4, 0));
            STEP(sp_apply_impl(s, SP_OP_FIND_PRESSURE, f_fp, 3, p_fp, 4, 0));
            STEP(sp_apply_impl(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0));
            STEP(sp_build_cells(s));
            STEP(sp_apply_impl(s, SP_OP_INTERNAL_FORCE, f_if, 6, p_if, 5, 0));
            STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
        } else
            return sp_fail(s, SP_ERR_INVALID, "unknown program id");
    }
#undef STEP
    return sp_time_end(s);
}
