// sp_sweep.cu — apply! / apply_binary! / apply_unary! on the device (reference src/core.jl:94-161).
//
// Gather formulation: one thread owns particle p (a slot of the cell-sorted SoA planes), scans the
// reference's 9/27 linear-offset cells, applies the exact distance predicate and accumulates in
// registers; nothing of q is written, so the sweep is race-free by construction (core.jl:122-123).
//
// Candidate cells are `key + dkey` with only the global range check 1 <= key <= key_max
// (core.jl:97-98) — including the row wrap-around at domain edges.  Cells with consecutive keys are
// contiguous in the sorted planes, so the three di = -1,0,1 cells of one (dj,dk) row form ONE
// contiguous slot range; the default order walks those 3 (2-D) / 9 (3-D) ranges, SP_FLAG_STRICT_ORDER
// walks the 9/27 cells in key_diff order (di outermost), which with the descending in-cell order is
// exactly the reference's accumulation order.
#include "sp_internal.cuh"
#include "sp_ops.cuh"

struct SweepCtx {
    const double *x, *y, *z;
    const int* cell_start;
    int n;
};

// Visit every candidate slot j of particle (xi,yi,zi): f(j, dx, dy, dz, d2).
template <bool STRICT, class F>
__device__ __forceinline__ void sp_for_candidates(const SpGrid& g, const SweepCtx& c, double xi, double yi, double zi,
                                                  F&& f) {
    const long long key = sp_find_key(g, xi, yi, zi);  // core.jl:95 recomputes the key from the current x
    const long long L1 = g.lim[0], L12 = g.lim[0] * g.lim[1];
    const int nk = (g.dim == 2) ? 1 : 3;
    if (STRICT) {
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
                for (int dk = (nk == 1 ? 0 : -1); dk <= (nk == 1 ? 0 : 1); dk++) {
                    const long long nkey = key + di + L1 * dj + L12 * dk;
                    if (nkey < 1 || nkey > g.key_max) continue;
                    const int jb = c.cell_start[nkey], je = c.cell_start[nkey + 1];
                    for (int j = jb; j < je; j++) {
                        double dx = __dsub_rn(xi, c.x[j]), dy = __dsub_rn(yi, c.y[j]), dz = __dsub_rn(zi, c.z[j]);
                        f(j, dx, dy, dz, sp_d2(dx, dy, dz));
                    }
                }
    } else {
        for (int dk = (nk == 1 ? 0 : -1); dk <= (nk == 1 ? 0 : 1); dk++)
            for (int dj = -1; dj <= 1; dj++) {
                const long long mid = key + L1 * dj + L12 * dk;
                long long klo = mid - 1, khi = mid + 1;
                if (klo < 1) klo = 1;
                if (khi > g.key_max) khi = g.key_max;
                if (klo > khi) continue;
                const int jb = c.cell_start[klo], je = c.cell_start[khi + 1];
                for (int j = jb; j < je; j++) {
                    double dx = __dsub_rn(xi, c.x[j]), dy = __dsub_rn(yi, c.y[j]), dz = __dsub_rn(zi, c.z[j]);
                    f(j, dx, dy, dz, sp_d2(dx, dy, dz));
                }
            }
    }
}

template <class Op, bool STRICT>
__global__ void __launch_bounds__(128) k_sweep(SpGrid g, SweepCtx c, typename Op::Params P, int self_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    if (!Op::active(P, i)) return;
    const double xi = c.x[i], yi = c.y[i], zi = c.z[i];
    typename Op::PS p;
    typename Op::Acc acc;
    Op::load(P, i, xi, yi, zi, p, acc);
    sp_for_candidates<STRICT>(g, c, xi, yi, zi, [&](int j, double dx, double dy, double dz, double d2) {
        // (r > h || p == q) && continue   (core.jl:105), r = sqrt_rn(d2)  <=>  d2 > T2
        if (d2 > g.T2 || j == i) return;
        Op::pair(P, p, j, dx, dy, dz, sqrt(d2), acc);
    });
    if (self_flag) Op::self(P, p, acc);
    Op::store(P, i, p, acc);
}

template <class U>
__global__ void __launch_bounds__(256) k_unary(typename U::Params P, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) U::apply(P, i);
}

template <class Op>
static int launch_sweep(sp_system* s, const typename Op::Params& P, int flags) {
    if (s->n == 0) return SP_OK;
    SweepCtx c;
    const double* X = s->fields[0].d;
    c.x = X;
    c.y = X + s->cap;
    c.z = X + 2 * s->cap;
    c.cell_start = s->cell_start;
    c.n = (int)s->n;
    const int self_flag = (flags & SP_FLAG_SELF) ? 1 : 0;
    if (flags & SP_FLAG_STRICT_ORDER)
        SP_LAUNCH(s, (k_sweep<Op, true>), sp_blocks(s->n, 128), 128, 0, s->g, c, P, self_flag);
    else
        SP_LAUNCH(s, (k_sweep<Op, false>), sp_blocks(s->n, 128), 128, 0, s->g, c, P, self_flag);
    return SP_OK;
}

template <template <class> class OpT, class MakeParams>
static int dispatch_kernel(sp_system* s, int kernel, double h, int flags, MakeParams&& mk) {
    SpKC kc;
    if (!sp_make_kc(kernel, h, &kc)) return sp_fail(s, SP_ERR_INVALID, "unknown SPH kernel id");
    switch (kernel) {
        case SP_KERNEL_WENDLAND1:
        case SP_KERNEL_WENDLAND2:
        case SP_KERNEL_WENDLAND3: {
            typename OpT<KWendland>::Params P;
            mk(P);
            P.kc = kc;
            return launch_sweep<OpT<KWendland>>(s, P, flags);
        }
        case SP_KERNEL_SPLINE23: {
            typename OpT<KSpline23>::Params P;
            mk(P);
            P.kc = kc;
            return launch_sweep<OpT<KSpline23>>(s, P, flags);
        }
        case SP_KERNEL_SPLINE24: {
            typename OpT<KSpline24>::Params P;
            mk(P);
            P.kc = kc;
            return launch_sweep<OpT<KSpline24>>(s, P, flags);
        }
    }
    return sp_fail(s, SP_ERR_INVALID, "unknown SPH kernel id");
}

template <class U>
static int launch_unary(sp_system* s, const typename U::Params& P) {
    if (s->n == 0) return SP_OK;
    SP_LAUNCH(s, (k_unary<U>), sp_blocks(s->n, 256), 256, 0, P, (int)s->n);
    return SP_OK;
}

static inline RV3 rv3(sp_system* s, int fid) {
    const double* d = s->fields[fid].d;
    return RV3{d, d + s->cap, d + 2 * s->cap};
}
static inline WV3 wv3(sp_system* s, int fid) {
    double* d = s->fields[fid].d;
    return WV3{d, d + s->cap, d + 2 * s->cap};
}
static inline double* sc(sp_system* s, int fid) { return s->fields[fid].d; }

// Internal entry shared with the step programs (sp_program.cu) and the CG (sp_isph.cu).
int sp_apply_impl(sp_system* s, int32_t op, const int32_t* F, int32_t nf, const double* Pm, int32_t np, int32_t flags) {
#define NEED(NC, NP, ...)                                                                    \
    {                                                                                        \
        const int _nc[] = {__VA_ARGS__};                                                     \
        int _rc = sp_check_fields(s, F, nf, _nc, NC);                                        \
        if (_rc) return _rc;                                                                 \
        if (np != (NP) || !Pm) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters for this operator"); \
    }
#define NEED_CELLS()                                                                                          \
    {                                                                                                         \
        if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "binary operator before sp_create_cell_list");    \
        if (F[0] != 0) return sp_fail(s, SP_ERR_INVALID, "the first field of a binary operator must be x (field 0)"); \
    }
    switch (op) {
        case SP_OP_BALANCE_OF_MASS: {
            NEED(4, 4, 3, 3, 1, 1);
            NEED_CELLS();
            return dispatch_kernel<OpBalanceOfMass>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.v = rv3(s, F[1]);
                P.rho = sc(s, F[2]);
                P.Drho = sc(s, F[3]);
                P.m = Pm[1];
                P.two_nu = Pm[3];
            });
        }
        case SP_OP_FIND_PRESSURE: {
            NEED(3, 4, 1, 1, 1);
            UFindPressure::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), Pm[0], Pm[1], Pm[2], Pm[3]};
            return launch_unary<UFindPressure>(s, P);
        }
        case SP_OP_INTERNAL_FORCE: {
            NEED(6, 5, 3, 3, 1, 1, 3, 1);
            NEED_CELLS();
            return dispatch_kernel<OpInternalForce>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.v = rv3(s, F[1]);
                P.P = sc(s, F[2]);
                P.rho = sc(s, F[3]);
                P.Dv = wv3(s, F[4]);
                P.type = sc(s, F[5]);
                P.m = Pm[1];
                P.visc = 2 * Pm[3] / (Pm[4] * Pm[4]);
            });
        }
        case SP_OP_INTERNAL_FORCE_CAVITY: {
            NEED(6, 6, 3, 3, 1, 1, 3, 1);
            NEED_CELLS();
            return dispatch_kernel<OpInternalForceCavity>(s, SP_KERNEL_WENDLAND2, Pm[1], flags, [&](auto& P) {
                P.v = rv3(s, F[1]);
                P.P = sc(s, F[2]);
                P.rho = sc(s, F[3]);
                P.Dv = wv3(s, F[4]);
                P.type = sc(s, F[5]);
                P.m = Pm[0];
                P.Re = Pm[2];
                P.vlid = Pm[3];
                P.ylid = Pm[4];
                P.lid = Pm[5];
                P.tenth_h = 0.1 * Pm[1];
                P.eps = 0.01 * (Pm[1] * Pm[1]);
            });
        }
        case SP_OP_MOVE: {
            NEED(4, 1, 3, 3, 3, 1);
            UMove::Params P{wv3(s, F[0]), rv3(s, F[1]), wv3(s, F[2]), sc(s, F[3]), Pm[0]};
            return launch_unary<UMove>(s, P);
        }
        case SP_OP_ACCELERATE: {
            NEED(3, 4, 3, 3, 1);
            UAccelerate::Params P{wv3(s, F[0]), rv3(s, F[1]), sc(s, F[2]), Pm[0], Pm[1], Pm[2], Pm[3]};
            return launch_unary<UAccelerate>(s, P);
        }
        case SP_OP_DENSITY_SUM: {
            NEED(2, 3, 3, 1);
            NEED_CELLS();
            return dispatch_kernel<OpDensitySum>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.out = sc(s, F[1]);
                P.m = Pm[1];
            });
        }
        case SP_OP_PRESSURE_FROM_RHO: {
            NEED(3, 1, 1, 1, 1);
            UPressureFromRho::Params P{sc(s, F[0]), sc(s, F[1]), sc(s, F[2]), Pm[0]};
            return launch_unary<UPressureFromRho>(s, P);
        }
        case SP_OP_INTERNAL_FORCE_SYM: {
            NEED(3, 4, 3, 1, 3);
            NEED_CELLS();
            return dispatch_kernel<OpInternalForceSym>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.P = sc(s, F[1]);
                P.a = wv3(s, F[2]);
                P.m = Pm[1];
                P.inv_rho0sq = 1.0 / (Pm[3] * Pm[3]);
            });
        }
        case SP_OP_FILL: {
            NEED(1, 1, 0);
            UFill::Params P{sc(s, F[0]), s->cap, s->fields[F[0]].ncomp, Pm[0]};
            return launch_unary<UFill>(s, P);
        }
        case SP_OP_ADVECT: {
            NEED(2, 1, 3, 3);
            UAdvect::Params P{wv3(s, F[0]), rv3(s, F[1]), Pm[0]};
            return launch_unary<UAdvect>(s, P);
        }
        case SP_OP_KICK: {
            NEED(2, 1, 3, 3);
            UKick::Params P{wv3(s, F[0]), rv3(s, F[1]), Pm[0]};
            return launch_unary<UKick>(s, P);
        }
        case SP_OP_ISPH_INITIALIZE: {
            NEED(6, 4, 3, 3, 1, 1, 1, 1);
            UIsphInitialize::Params P{wv3(s, F[0]), wv3(s, F[1]), sc(s, F[2]), sc(s, F[3]), sc(s, F[4]),
                                      sc(s, F[5]),  Pm[0],        Pm[1],       Pm[2],       Pm[3]};
            return launch_unary<UIsphInitialize>(s, P);
        }
        case SP_OP_ISPH_VISCOUS_FORCE: {
            NEED(3, 5, 3, 3, 3);
            NEED_CELLS();
            return dispatch_kernel<OpIsphViscous>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.v = rv3(s, F[1]);
                P.Dv = wv3(s, F[2]);
                P.coef = 2.0 * Pm[1] * Pm[3] / (Pm[4] * Pm[4]);
            });
        }
        case SP_OP_ISPH_DIV_L_LAMBDA: {
            NEED(5, 5, 3, 3, 1, 1, 1);
            NEED_CELLS();
            return dispatch_kernel<OpIsphDivLLambda>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.v = rv3(s, F[1]);
                P.div = sc(s, F[2]);
                P.L = sc(s, F[3]);
                P.lambda = sc(s, F[4]);
                P.m = Pm[1];
                P.m_over_rho = Pm[1] / Pm[3];
                P.inv_dim = 1.0 / Pm[4];
            });
        }
        case SP_OP_ISPH_PROJECTION_VECTOR: {
            NEED(2, 2, 1, 1);
            UIsphProjectionVector::Params P{sc(s, F[0]), sc(s, F[1]), -(Pm[0] * Pm[0]), Pm[1]};
            return launch_unary<UIsphProjectionVector>(s, P);
        }
        case SP_OP_ISPH_INTERNAL_FORCE: {
            NEED(3, 4, 3, 1, 3);
            NEED_CELLS();
            return dispatch_kernel<OpIsphInternalForce>(s, (int)Pm[0], Pm[2], flags, [&](auto& P) {
                P.P = sc(s, F[1]);
                P.Dv = wv3(s, F[2]);
                P.coef = Pm[1] / (Pm[3] * Pm[3]);
            });
        }
        case SP_OP_ISPH_ACCELERATE: {
            NEED(3, 1, 3, 3, 1);
            UIsphAccelerate::Params P{wv3(s, F[0]), wv3(s, F[1]), sc(s, F[2]), Pm[0]};
            return launch_unary<UIsphAccelerate>(s, P);
        }
    }
    return sp_fail(s, SP_ERR_INVALID, "unknown operator id");
}

// fields {x, L, lambda, type, p_in, y_out}; params {kernel, m, h, rho, C_free}
int sp_poisson_apply_impl(sp_system* s, const int32_t* F, int32_t nf, const double* Pm, int32_t np) {
    const int32_t op = 0;
    (void)op;
    NEED(6, 5, 3, 1, 1, 1, 1, 1);
    NEED_CELLS();
    return dispatch_kernel<OpPoissonApply>(s, (int)Pm[0], Pm[2], 0, [&](auto& P) {
        P.L = sc(s, F[1]);
        P.lambda = sc(s, F[2]);
        P.type = sc(s, F[3]);
        P.pin = sc(s, F[4]);
        P.y = sc(s, F[5]);
        P.off_coef = 2.0 * (Pm[2] * Pm[2]) * Pm[1] / Pm[3];
        P.h2 = Pm[2] * Pm[2];
        P.C_free = Pm[4];
    });
}
#undef NEED
#undef NEED_CELLS

// ------------------------------------------------------------------ neighbour-list export (parity view)
__global__ void k_nbr_count(SpGrid g, SweepCtx c, const int* ref, long long* counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    int cnt = 0;
    sp_for_candidates<true>(g, c, c.x[i], c.y[i], c.z[i], [&](int j, double, double, double, double d2) {
        if (d2 > g.T2 || j == i) return;
        cnt++;
    });
    counts[ref[i]] = cnt;
}
__global__ void k_nbr_fill(SpGrid g, SweepCtx c, const int* ref, const long long* offsets, long long* ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    long long o = offsets[ref[i]];
    sp_for_candidates<true>(g, c, c.x[i], c.y[i], c.z[i], [&](int j, double, double, double, double d2) {
        if (d2 > g.T2 || j == i) return;
        ids[o++] = (long long)ref[j] + 1;
    });
}

extern "C" {

int32_t sp_apply(sp_system* s, int32_t op, const int32_t* fields, int32_t nfields, const double* params, int32_t nparams,
                 int32_t flags) {
    if (!s) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_time_begin(s);
    if (rc) return rc;
    if ((rc = sp_apply_impl(s, op, fields, nfields, params, nparams, flags))) return rc;
    return sp_time_end(s);
}

int32_t sp_poisson_apply(sp_system* s, const int32_t* fields, int32_t nfields, const double* params, int32_t nparams) {
    if (!s) return SP_ERR_INVALID;
    SP_CUDA(s, cudaSetDevice(s->device));
    int rc = sp_time_begin(s);
    if (rc) return rc;
    if ((rc = sp_poisson_apply_impl(s, fields, nfields, params, nparams))) return rc;
    return sp_time_end(s);
}

int32_t sp_get_neighbour_lists(sp_system* s, int64_t* offsets, int64_t* ids, int64_t ids_cap) {
    if (!s || !offsets) return SP_ERR_INVALID;
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "no cell list: call sp_create_cell_list first");
    SP_CUDA(s, cudaSetDevice(s->device));
    const long long n = s->n;
    offsets[0] = 0;
    if (n == 0) return SP_OK;
    SweepCtx c;
    const double* X = s->fields[0].d;
    c.x = X;
    c.y = X + s->cap;
    c.z = X + 2 * s->cap;
    c.cell_start = s->cell_start;
    c.n = (int)n;
    int rc = sp_ensure_stage(s, n + 1);
    if (rc) return rc;
    long long* counts = (long long*)s->stage;
    SP_LAUNCH(s, k_nbr_count, sp_blocks(n, 128), 128, 0, s->g, c, s->ref, counts);
    std::vector<long long> h(n + 1);
    SP_CUDA(s, cudaMemcpyAsync(h.data(), counts, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    long long run = 0;
    for (long long i = 0; i < n; i++) {
        long long cnt = h[i];
        h[i] = run;
        offsets[i] = run;
        run += cnt;
    }
    h[n] = run;
    offsets[n] = run;
    if (!ids) return SP_OK;
    if (ids_cap < run) return sp_fail(s, SP_ERR_INVALID, "ids buffer too small");
    if (run == 0) return SP_OK;
    long long *d_off = nullptr, *d_ids = nullptr;
    SP_CUDA(s, cudaMalloc(&d_off, (size_t)(n + 1) * sizeof(long long)));
    cudaError_t e = cudaMalloc(&d_ids, (size_t)run * sizeof(long long));
    if (e != cudaSuccess) {
        cudaFree(d_off);
        return sp_fail_cuda(s, e, "cudaMalloc ids", __FILE__, __LINE__);
    }
    cudaMemcpyAsync(d_off, h.data(), (size_t)(n + 1) * sizeof(long long), cudaMemcpyHostToDevice, s->stream);
    k_nbr_fill<<<sp_blocks(n, 128), 128, 0, s->stream>>>(s->g, c, s->ref, d_off, d_ids);
    s->launches++;
    cudaMemcpyAsync(ids, d_ids, (size_t)run * sizeof(long long), cudaMemcpyDeviceToHost, s->stream);
    e = cudaStreamSynchronize(s->stream);
    cudaFree(d_off);
    cudaFree(d_ids);
    if (e != cudaSuccess) return sp_fail_cuda(s, e, "neighbour list export", __FILE__, __LINE__);
    return SP_OK;
}

}  // extern "C"
