// sp_isph.cu — ISPH pressure solve: un-preconditioned CG on the matrix-free operator of sp_sweep.cu.
//
// Replaces `A = assemble_matrix(sys, projection_matrix); P .= cg(A, b)` of
// examples/collapse_dry_implicit.jl:223-227 (assemble_matrix: src/core.jl:196-225).  The reference
// assembles a SparseMatrixCSC serially; here A is never formed.  cg is IterativeSolvers.jl (third-party,
// unpinned): x0 = 0, tol = max(reltol*|b|, abstol), maxiter = N, iteration
//   beta = |r|^2/|r_prev|^2; u = r + beta u; c = A u; alpha = |r|^2/(u.c); x += alpha u; r -= alpha c.
#include <cmath>

#include "sp_internal.cuh"

int sp_poisson_apply_impl(sp_system* s, const int32_t* F, int32_t nf, const double* Pm, int32_t np);
int sp_dot_device(sp_system* s, const double* a, const double* b, long long n, double* partial, double* out3);
int sp_slab_allreduce_device(sp_system* s, double* d_inout, int count, int is_max);  // no-op without a slab

__global__ void k_cg_init(const double* b, double* r, double* u, double* x, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) {
        r[i] = b[i];
        u[i] = 0.0;
        x[i] = 0.0;
    }
}
__global__ void k_cg_dir(const double* r, double* u, double beta, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) u[i] = r[i] + beta * u[i];
}
// alpha = res2 / (u.c) with u.c read from device memory
__global__ void k_cg_step(double* x, double* r, const double* u, const double* c, double res2, const double* uc,
                          long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) {
        const double alpha = res2 / uc[0];
        x[i] += alpha * u[i];
        r[i] -= alpha * c[i];
    }
}

static int scratch_field(sp_system* s, const char* name, int32_t* fid) {
    int rc = sp_add_field(s, name, 1, fid);
    if (rc) return rc;
    s->fields[*fid].transient = true;
    return SP_OK;
}

extern "C" int32_t sp_poisson_cg(sp_system* s, const int32_t* F, int32_t nf, const double* Pm, int32_t np, double reltol,
                                 double abstol, int64_t maxiter, int64_t* iters, double* resid) {
    if (!s) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    const int nc[] = {3, 1, 1, 1, 1, 1};
    int rc = sp_check_fields(s, F, nf, nc, 6);
    if (rc) return rc;
    if (np != 5 || !Pm) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters");
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "sp_poisson_cg before sp_create_cell_list");
    const long long n = s->n;
    if (iters) *iters = 0;
    if (resid) *resid = 0.0;
    if (n == 0) return SP_OK;
    int32_t fr, fu, fc;
    if ((rc = scratch_field(s, "_cg_r", &fr)) || (rc = scratch_field(s, "_cg_u", &fu)) ||
        (rc = scratch_field(s, "_cg_c", &fc)))
        return rc;
    if (!s->dscal) {
        SP_CUDA(s, sp_dmalloc(&s->dscal, (3 * 1024 + 16) * sizeof(double)));
        SP_CUDA(s, cudaHostAlloc(&s->h_scal, 16 * sizeof(double), cudaHostAllocDefault));
    }
    if ((rc = sp_time_begin(s))) return rc;
    double* partial = s->dscal;
    double* d_rr = s->dscal + 3 * 1024;      // |r|^2
    double* d_uc = s->dscal + 3 * 1024 + 4;  // u.c
    double* b = s->fields[F[4]].d;
    double* x = s->fields[F[5]].d;
    double *r = s->fields[fr].d, *u = s->fields[fu].d, *c = s->fields[fc].d;
    const int B = 256;
    const int32_t Fa[6] = {F[0], F[1], F[2], F[3], fu, fc};
    sp_wrote(s, F[5]);
    sp_wrote(s, fr);
    sp_wrote(s, fu);
    sp_wrote(s, fc);
    SP_LAUNCH(s, k_cg_init, sp_blocks(n, B), B, 0, b, r, u, x, n);
    auto norm2 = [&](double* out_host) -> int {
        int rc2 = sp_dot_device(s, r, r, n, partial, d_rr);  // ghosts are masked inside
        if (rc2) return rc2;
        if ((rc2 = sp_slab_allreduce_device(s, d_rr, 1, 0))) return rc2;
        SP_CUDA(s, cudaMemcpyAsync(s->h_scal, d_rr, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        SP_CUDA(s, cudaStreamSynchronize(s->stream));
        *out_host = s->h_scal[0];
        return SP_OK;
    };
    double res2;
    if ((rc = norm2(&res2))) return rc;
    double residual = std::sqrt(res2), prev = 1.0;
    const double tol = std::fmax(reltol * residual, abstol);
    if (maxiter <= 0) maxiter = n;
    int64_t it = 0;
    while (it < maxiter && residual > tol) {
        const double beta = residual * residual / (prev * prev);
        SP_LAUNCH(s, k_cg_dir, sp_blocks(n, B), B, 0, r, u, beta, n);
        if (s->slab) {  // ghosts of the search direction come from their owners
            const int32_t fh[1] = {fu};
            if ((rc = sp_slab_halo_refresh(s, fh, 1))) return rc;
        }
        if ((rc = sp_poisson_apply_impl(s, Fa, 6, Pm, 5))) return rc;
        if ((rc = sp_dot_device(s, u, c, n, partial, d_uc))) return rc;
        if ((rc = sp_slab_allreduce_device(s, d_uc, 1, 0))) return rc;
        SP_LAUNCH(s, k_cg_step, sp_blocks(n, B), B, 0, x, r, u, c, residual * residual, d_uc, n);
        prev = residual;
        if ((rc = norm2(&res2))) return rc;
        residual = std::sqrt(res2);
        it++;
    }
    if ((rc = sp_time_end(s))) return rc;
    if (iters) *iters = it;
    if (resid) *resid = residual;
    if (!(residual <= tol)) return sp_fail(s, SP_ERR_NOT_CONVERGED, "CG reached maxiter before the tolerance");
    return SP_OK;
}
