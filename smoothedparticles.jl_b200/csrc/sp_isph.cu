// sp_isph.cu — ISPH pressure solve: un-preconditioned CG on the matrix-free operator of sp_sweep.cu.
//
// Replaces `A = assemble_matrix(sys, projection_matrix); P .= cg(A, b)` of
// examples/collapse_dry_implicit.jl:223-227 (assemble_matrix: src/core.jl:196-225).  The reference
// assembles a SparseMatrixCSC serially; here A is never formed.  cg is IterativeSolvers.jl (third-party,
// unpinned): x0 = 0, tol = max(reltol*|b|, abstol), maxiter = N, iteration
//   beta = |r|^2/|r_prev|^2; u = r + beta u; c = A u; alpha = |r|^2/(u.c); x += alpha u; r -= alpha c.
#include <cooperative_groups.h>

#include <cmath>
#include <cstdlib>

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

// ------------------------------------------------------------------ persistent CG (single-GPU systems)
// The whole solve is ONE cooperative kernel: the operator's coefficients sit in the ELL layout of the cached
// neighbour lists (sp_poisson_ell_build), each iteration is  u = r + beta u | c = A u, u.c | x += alpha u,
// r -= alpha c, r.r  separated by three grid syncs, the dot products are reduced deterministically (per-CTA partial
// in a fixed tree, then every CTA adds the partials in the same order), and the convergence test runs on the
// device: no launch, no host round trip per iteration (the host-driven loop below costs ~45 us per iteration on
// the 23 k-particle ISPH config, this one ~8 us).
namespace cg = cooperative_groups;

struct CgArgs {
    long long n;
    const int *cnt, *ids;
    const double *aval, *diag, *b;
    const double* ghost;  // unused here (slab systems take the host-driven path)
    double *x, *r, *u, *c;
    double* partial;  // 2 x 1024 doubles
    double* out;      // [0] iterations, [1] residual, [2] tol
    long long maxiter;
    double reltol, abstol;
    int capk;
};

__device__ __forceinline__ double cg_block_sum(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();  // sh may still be read from the previous call
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = blockDim.x >> 5;
    for (int k = 0; k < nw; k++) t += sh[k];  // same order in every thread
    return t;
}

// deterministic grid-wide sum, returned to every thread of every CTA; `buf` alternates between two partial arrays
__device__ __forceinline__ double cg_grid_sum(cg::grid_group& grid, double v, double* partial, int& buf, double* sh) {
    double* p = partial + 1024 * buf;
    buf ^= 1;
    const double bs = cg_block_sum(v, sh);
    if (threadIdx.x == 0) p[blockIdx.x] = bs;
    grid.sync();
    double t = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) t += p[k];
    return cg_block_sum(t, sh);
}

__global__ void __launch_bounds__(256) k_cg_persistent(CgArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[8];
    int buf = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double local = 0.0;
    for (long long i = t0; i < a.n; i += stride) {
        const double bi = a.b[i];
        a.r[i] = bi;
        a.u[i] = 0.0;
        a.x[i] = 0.0;
        local += bi * bi;
    }
    double res2 = cg_grid_sum(grid, local, a.partial, buf, sh);
    double residual = sqrt(res2), prev = 1.0;
    const double tol = fmax(a.reltol * residual, a.abstol);
    long long it = 0;
    while (it < a.maxiter && residual > tol) {
        const double beta = residual * residual / (prev * prev);
        for (long long i = t0; i < a.n; i += stride) a.u[i] = a.r[i] + beta * a.u[i];
        grid.sync();  // the mat-vec gathers u of other rows
        local = 0.0;
        for (long long i = t0; i < a.n; i += stride) {
            const int n_nb = a.cnt[i];
            const size_t base = ((size_t)(i >> 5) * a.capk << 5) + (i & 31);
            double sum = 0.0;
            for (int k = 0; k < n_nb; k++) sum += a.aval[base + ((size_t)k << 5)] * a.u[a.ids[base + ((size_t)k << 5)]];
            const double ui = a.u[i];
            const double ci = sum + a.diag[i] * ui;
            a.c[i] = ci;
            local += ui * ci;
        }
        const double uc = cg_grid_sum(grid, local, a.partial, buf, sh);
        const double alpha = residual * residual / uc;
        local = 0.0;
        for (long long i = t0; i < a.n; i += stride) {
            a.x[i] += alpha * a.u[i];
            const double ri = a.r[i] - alpha * a.c[i];
            a.r[i] = ri;
            local += ri * ri;
        }
        prev = residual;
        res2 = cg_grid_sum(grid, local, a.partial, buf, sh);
        residual = sqrt(res2);
        it++;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.out[0] = (double)it;
        a.out[1] = residual;
        a.out[2] = tol;
    }
}

int sp_poisson_ell_build(sp_system* s, const int32_t* F, const double* Pm, double* aval, double* diag, int* d_overflow,
                         const int** ids_out, const int** cnt_out);
int sp_nbr_prepare(sp_system* s, int* capk);

static int scratch_field(sp_system* s, const char* name, int32_t* fid) {
    int rc = sp_add_field(s, name, 1, fid);
    if (rc) return rc;
    s->fields[*fid].transient = true;
    return SP_OK;
}

static int cg_persistent(sp_system* s, const int32_t* F, const double* Pm, int32_t fr, int32_t fu, int32_t fc, double reltol,
                         double abstol, int64_t maxiter, int64_t* iters, double* resid, int* done) {
    *done = 0;
    const long long n = s->n;
    int capk = 0;
    int rc = sp_nbr_prepare(s, &capk);  // the lists (and their capacity) of the current positions
    if (rc) return rc;
    const size_t need = (size_t)s->cap * capk * sizeof(double);
    if (!s->ell_val || s->ell_cap != s->cap || s->ell_capk != capk) {
        size_t free_b = 0, total_b = 0;
        SP_CUDA(s, cudaMemGetInfo(&free_b, &total_b));
        if (s->ell_val) SP_CUDA(s, sp_dfree(s, s->ell_val));
        s->ell_val = nullptr;
        s->ell_cap = 0;
        if (need > free_b / 2) return SP_OK;  // too large: matrix-free path
        SP_CUDA(s, sp_dmalloc(&s->ell_val, need));
        s->ell_cap = s->cap;
        s->ell_capk = capk;
    }
    int32_t fdiag;
    rc = sp_add_field(s, "_cg_diag", 1, &fdiag);
    if (rc) return rc;
    s->fields[fdiag].transient = true;
    int* d_flag = s->counters + 32;
    SP_CUDA(s, cudaMemsetAsync(d_flag, 0, sizeof(int), s->stream));
    const int *ids = nullptr, *cnt = nullptr;
    if ((rc = sp_poisson_ell_build(s, F, Pm, s->ell_val, s->fields[fdiag].d, d_flag, &ids, &cnt))) return rc;
    SP_CUDA(s, cudaMemcpyAsync(s->h_counters + 32, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    if (s->h_counters[32]) return SP_OK;  // a list overflowed: matrix-free path
    static int blocks_per_sm = 0, n_sm = 0;
    if (!blocks_per_sm) {
        SP_CUDA(s, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_cg_persistent, 256, 0));
        SP_CUDA(s, cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device));
    }
    long long grid = std::min<long long>((n + 255) / 256, (long long)blocks_per_sm * n_sm);
    grid = std::min<long long>(std::max<long long>(grid, 1), 1024);
    CgArgs a;
    a.n = n;
    a.cnt = cnt;
    a.ids = ids;
    a.aval = s->ell_val;
    a.diag = s->fields[fdiag].d;
    a.b = s->fields[F[4]].d;
    a.ghost = nullptr;
    a.x = s->fields[F[5]].d;
    a.r = s->fields[fr].d;
    a.u = s->fields[fu].d;
    a.c = s->fields[fc].d;
    a.partial = s->dscal;
    a.out = s->dscal + 3 * 1024;
    a.maxiter = maxiter > 0 ? maxiter : n;
    a.reltol = reltol;
    a.abstol = abstol;
    a.capk = capk;
    sp_wrote(s, F[5]);
    sp_wrote(s, fr);
    sp_wrote(s, fu);
    sp_wrote(s, fc);
    void* args[] = {&a};
    SP_CUDA(s, cudaLaunchCooperativeKernel((void*)k_cg_persistent, dim3((unsigned)grid), dim3(256), args, 0, s->stream));
    s->launches++;
    SP_CUDA(s, cudaMemcpyAsync(s->h_scal, a.out, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    if (iters) *iters = (int64_t)s->h_scal[0];
    if (resid) *resid = s->h_scal[1];
    *done = 1;
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
    if ((rc = sp_settle(s))) return rc;
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
    // single-GPU systems: the persistent cooperative kernel (falls through to the host-driven loop when a target
    // has more neighbours than the lists hold, when the coefficient array would not fit, or with SP_CG_PERSISTENT=0)
    if (!s->slab && !(getenv("SP_CG_PERSISTENT") && atoi(getenv("SP_CG_PERSISTENT")) == 0)) {
        int done = 0;
        if ((rc = cg_persistent(s, F, Pm, fr, fu, fc, reltol, abstol, maxiter, iters, resid, &done))) return rc;
        if (done) {
            if ((rc = sp_time_end(s))) return rc;
            if (!(s->h_scal[1] <= s->h_scal[2]))  // residual, tol as the kernel left them
                return sp_fail(s, SP_ERR_NOT_CONVERGED, "CG reached maxiter before the tolerance");
            return SP_OK;
        }
    }
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
    if (s->slab) {  // the ghost copies of the solution come from their owners (internal_force! reads them next)
        const int32_t fp[1] = {F[5]};
        if ((rc = sp_slab_halo_refresh(s, fp, 1))) return rc;
    }
    if ((rc = sp_time_end(s))) return rc;
    if (iters) *iters = it;
    if (resid) *resid = residual;
    if (!(residual <= tol)) return sp_fail(s, SP_ERR_NOT_CONVERGED, "CG reached maxiter before the tolerance");
    return SP_OK;
}

// ------------------------------------------------------------------ assemble_matrix as COO triplets
__global__ void __launch_bounds__(256) k_coo_counts(const int* __restrict__ cnt, int* __restrict__ off, long long n, int mult) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) off[i] = cnt[i] + mult;  // the neighbours and the diagonal (once per visit of the own cell)
}
__global__ void __launch_bounds__(128) k_coo_export(long long n, int capk, const int* __restrict__ cnt,
                                                    const int* __restrict__ ids, const double* __restrict__ aval,
                                                    const double* __restrict__ diag, const int* __restrict__ ref,
                                                    const int* __restrict__ off, long long* __restrict__ I,
                                                    long long* __restrict__ J, double* __restrict__ V, int mult) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long o = off[i];
    const long long row = (long long)ref[i] + 1;  // 1-based index in the reference's sys.particles
    for (int v = 0; v < mult; v++, o++) {  // diag[i] is the sum over the visits; the reference pushes one triplet per visit
        I[o] = row;
        J[o] = row;
        V[o] = diag[i] / (double)mult;
    }
    const size_t base = ((size_t)(i >> 5) * capk << 5) + (i & 31);
    const int n_nb = cnt[i];
    for (int k = 0; k < n_nb; k++, o++) {
        const size_t e = base + ((size_t)k << 5);
        I[o] = row;
        J[o] = (long long)ref[ids[e]] + 1;
        V[o] = aval[e];
    }
}

extern "C" int32_t sp_assemble_matrix(sp_system* s, const int32_t* F, int32_t nf, const double* Pm, int32_t np, int64_t* I,
                                      int64_t* J, double* V, int64_t cap, int64_t* nnz) {
    if (!s || !nnz) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    const int nc[] = {3, 1, 1, 1};
    int rc = sp_check_fields(s, F, nf, nc, 4);
    if (rc) return rc;
    if (np != 5 || !Pm) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters");
    if (F[0] != 0) return sp_fail(s, SP_ERR_INVALID, "the first field must be x (field 0)");
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "sp_assemble_matrix before sp_create_cell_list");
    if (s->slab) return sp_fail(s, SP_ERR_STATE, "sp_assemble_matrix is not available on a slab system");
    const bool fill = I || J || V;
    if (fill && !(I && J && V)) return SP_ERR_INVALID;
    *nnz = 0;
    if ((rc = sp_settle(s))) return rc;
    const long long n = s->n;
    if (n == 0) return SP_OK;
    int capk = 0;
    if ((rc = sp_nbr_prepare(s, &capk))) return rc;
    int *d_off = nullptr, *d_flag = s->counters + 32;
    double *aval = nullptr, *diag = nullptr;
    long long *dI = nullptr, *dJ = nullptr;
    double* dV = nullptr;
    auto release = [&]() {
        sp_dfree(s, d_off);
        sp_dfree(s, aval);
        sp_dfree(s, diag);
        sp_dfree(s, dI);
        sp_dfree(s, dJ);
        sp_dfree(s, dV);
    };
#define SP_TRY(call)                                                                   \
    do {                                                                               \
        cudaError_t _e = (call);                                                       \
        if (_e != cudaSuccess) {                                                       \
            release();                                                                 \
            return sp_fail_cuda(s, _e, #call, __FILE__, __LINE__);                     \
        }                                                                              \
    } while (0)
    SP_TRY(sp_dmalloc(&d_off, (size_t)(n + 1) * sizeof(int)));
    {
        k_coo_counts<<<sp_blocks(n, 256), 256, 0, s->stream>>>(s->nbr_cnt, d_off, n, sp_self_visits(s));
        s->launches++;
        SP_TRY(cudaGetLastError());
    }
    // total = last offset + last count, read before the arrays are filled (two-call protocol)
    int last_cnt = 0, last_off = 0;
    SP_TRY(cudaMemcpyAsync(&last_cnt, d_off + n - 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_TRY(cudaStreamSynchronize(s->stream));
    if ((rc = sp_exclusive_scan_i32(s, d_off, n))) {
        release();
        return rc;
    }
    SP_TRY(cudaMemcpyAsync(&last_off, d_off + n - 1, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_TRY(cudaStreamSynchronize(s->stream));
    const long long total = (long long)last_off + last_cnt;
    if (total < 0 || total > 2000000000LL) {
        release();
        return sp_fail(s, SP_ERR_INVALID, "assemble_matrix: more than 2e9 triplets");
    }
    *nnz = total;
    if (!fill) {
        release();
        return SP_OK;
    }
    if (cap < total) {
        release();
        return sp_fail(s, SP_ERR_INVALID, "assemble_matrix: triplet arrays too small");
    }
    SP_TRY(sp_dmalloc(&aval, (size_t)s->cap * capk * sizeof(double)));
    SP_TRY(sp_dmalloc(&diag, (size_t)n * sizeof(double)));
    SP_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), s->stream));
    const int *ids = nullptr, *cnt = nullptr;
    if ((rc = sp_poisson_ell_build(s, F, Pm, aval, diag, d_flag, &ids, &cnt))) {
        release();
        return rc;
    }
    int overflow = 0;
    SP_TRY(cudaMemcpyAsync(&overflow, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    SP_TRY(cudaStreamSynchronize(s->stream));
    if (overflow) {  // cannot happen after sp_nbr_prepare has grown the lists; kept as a loud guard
        release();
        return sp_fail(s, SP_ERR_STATE, "assemble_matrix: a neighbour list overflowed its capacity");
    }
    SP_TRY(sp_dmalloc(&dI, (size_t)total * sizeof(long long)));
    SP_TRY(sp_dmalloc(&dJ, (size_t)total * sizeof(long long)));
    SP_TRY(sp_dmalloc(&dV, (size_t)total * sizeof(double)));
    {
        k_coo_export<<<sp_blocks(n, 128), 128, 0, s->stream>>>(n, capk, cnt, ids, aval, diag, s->ref, d_off, dI, dJ, dV, sp_self_visits(s));
        s->launches++;
        SP_TRY(cudaGetLastError());
    }
    SP_TRY(cudaMemcpyAsync(I, dI, (size_t)total * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    SP_TRY(cudaMemcpyAsync(J, dJ, (size_t)total * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    SP_TRY(cudaMemcpyAsync(V, dV, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    SP_TRY(cudaStreamSynchronize(s->stream));
#undef SP_TRY
    release();
    return SP_OK;
}
