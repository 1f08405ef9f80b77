// sp_program.cu — fused step programs: the time loops of the examples issued from inside the library,
// so a Julia/Python host pays one FFI crossing per batch of steps instead of 7-9 per step.
// Same kernels, same order, same arithmetic as the per-call path.
#include "sp_internal.cuh"

int sp_apply_impl(sp_system* s, int32_t op, const int32_t* F, int32_t nf, const double* Pm, int32_t np, int32_t flags);
int sp_build_cells(sp_system* s);
int sp_kick_kick_move_impl(sp_system* s, const int32_t* F, const double* Pm);
int sp_find_pressure_pr_impl(sp_system* s, const int32_t* F, const double* Pm);

// fields {x, v, Dv, rho, Drho, P, type}; params {kernel, m, h, two_nu, dt, c2, rho0, mu, gx, gy, gz}
static int run_program(sp_system* s, int32_t program, const int32_t* F, int32_t nf, const double* P, int32_t np,
                       int64_t nsteps);

extern "C" int32_t sp_run_program(sp_system* s, int32_t program, const int32_t* F, int32_t nf, const double* P,
                                  int32_t np, int64_t nsteps) {
    if (!s) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_time_begin(s);
    if (rc) return rc;
    s->in_program = true;
    rc = run_program(s, program, F, nf, P, np, nsteps);
    s->in_program = false;
    if (rc) return rc;
    return sp_time_end(s);
}

// On a slab system (sp_slab.cu) the cell-list build is the slab rebuild (migration + ghost halos + local build) and
// the ghost copies of rho and P are refreshed after find_pressure! (ghosts cannot integrate their own Drho).
static int run_program(sp_system* s, int32_t program, const int32_t* F, int32_t nf, const double* P, int32_t np,
                       int64_t nsteps) {
    const int nc[] = {3, 3, 3, 1, 1, 1, 1};
    int rc = sp_check_fields(s, F, nf, nc, 7);
    if (rc) return rc;
    if (np != 11 || !P) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters for this program");
    const bool slab = s->slab != nullptr;
    if (slab && program != SP_PROGRAM_WCSPH_3D)
        return sp_fail(s, SP_ERR_STATE, "only the 3-D WCSPH program runs on slab systems");
    const int32_t x = F[0], v = F[1], Dv = F[2], rho = F[3], Drho = F[4], Pr = F[5], ty = F[6];
    const double kernel = P[0], m = P[1], h = P[2], two_nu = P[3], dt = P[4], c2 = P[5], rho0 = P[6], mu = P[7];
    const int32_t f_bom[4] = {x, v, rho, Drho}, f_fp[3] = {rho, Drho, Pr}, f_if[6] = {x, v, Pr, rho, Dv, ty},
                  f_mv[4] = {x, v, Dv, ty}, f_ac[3] = {v, Dv, ty};
    const double p_bom[4] = {kernel, m, h, two_nu}, p_fp[4] = {dt, c2, rho0, 0.0}, p_if[5] = {kernel, m, h, mu, rho0},
                 p_ac[4] = {0.5 * dt, P[8], P[9], P[10]};
    const int32_t f_halo[2] = {rho, Pr};
#define STEP(call) \
    if ((rc = (call))) return rc;
    for (int64_t k = 0; k < nsteps; k++) {
        if (program == SP_PROGRAM_WCSPH_3D) {  // examples/collapse3d.jl:136-150
            // same statements in the same order as the per-call loop; the unary passes that touch the same
            // fields back to back are issued as one kernel each (bit-identical, tested against the per-call path):
            //   accelerate!, accelerate! of step k with move! of step k+1;  find_pressure! with the P/rho^2 pass
            const double p_mv[1] = {dt};
            const int32_t f_kkm[4] = {v, Dv, x, ty};
            const double p_kkm[5] = {0.5 * dt, P[8], P[9], P[10], dt};
            if (k == 0) STEP(sp_apply_impl(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0));
            if (slab) {
                STEP(sp_slab_create_cell_list(s));
            } else {
                STEP(sp_build_cells(s));
            }
            STEP(sp_apply_impl(s, SP_OP_BALANCE_OF_MASS, f_bom, 4, p_bom, 4, 0));
            if (slab) {
                // the ghosts' rho and P come from their owners, so P/rho^2 is evaluated after the refresh
                STEP(sp_apply_impl(s, SP_OP_FIND_PRESSURE, f_fp, 3, p_fp, 4, 0));
                STEP(sp_slab_halo_refresh(s, f_halo, 2));
                STEP(sp_apply_impl(s, SP_OP_INTERNAL_FORCE, f_if, 6, p_if, 5, 0));
            } else {
                STEP(sp_find_pressure_pr_impl(s, f_fp, p_fp));
                STEP(sp_apply_impl(s, SP_OP_INTERNAL_FORCE, f_if, 6, p_if, 5, SP_FLAG_INTERNAL_PR_READY));
            }
            if (k + 1 < nsteps) {
                STEP(sp_kick_kick_move_impl(s, f_kkm, p_kkm));
            } else {
                STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
                STEP(sp_apply_impl(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0));
            }
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
    return SP_OK;
}
