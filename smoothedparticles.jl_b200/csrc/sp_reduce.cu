// sp_reduce.cu — diagnostics reductions, point sums and kernel-function evaluation.
//   sp_reduce         the serial energy / front loops of the examples (collapse_dry.jl:166-187,
//                     test_collision_2d.jl:96-100, collapse_dry_implicit.jl:173-177)
//   sp_sum_at_points  SmoothedParticles.sum(sys, f, x)  (src/core.jl:240-260; cavity_flow.jl:162-180)
//   sp_kernel_eval    src/kernels.jl on the device (port of tests/test_kernels.jl runs against it)
// Reductions are two-stage (per-block partials, then one block) so the result is deterministic.
#include "sp_internal.cuh"
#include "sp_ops.cuh"

int sp_slab_allreduce_device(sp_system* s, double* d_inout, int count, int is_max);  // sp_slab.cu

#define RED_B 256
#define RED_MAXBLOCKS 1024

struct RedParams {
    const double* f[4];  // field bases
    const double* ghost; // slab systems: non-zero = ghost copy, skipped (nullptr otherwise)
    long long cap;
    double p[8];
    int ncomp;
};

template <int RED>
__device__ __forceinline__ void red_map(const RedParams& R, long long i, double v[3]) {
    const long long cap = R.cap;
    if (RED == SP_RED_ENERGY_WCSPH) {  // fields {x, v, rho}; params {m, c, rho0, gx, gy, gz}
        const double *X = R.f[0], *V = R.f[1];
        double m = R.p[0], c = R.p[1], rho0 = R.p[2];
        double vx = V[i], vy = V[cap + i], vz = V[2 * cap + i];
        double kinetic = 0.5 * m * (vx * vx + vy * vy + vz * vz);
        double potential = -m * (R.p[3] * X[i] + R.p[4] * X[cap + i] + R.p[5] * X[2 * cap + i]);
        double rho = R.f[2][i];
        double internal = m * (c * c) * (log(fabs(rho / rho0)) + rho0 / rho - 1.0);
        v[0] = kinetic + potential + internal;
    } else if (RED == SP_RED_FRONT) {  // fields {x, type}; params {width, height, h, xmax}
        const double* X = R.f[0];
        double t = R.f[1][i];
        double x1 = X[i], x2 = X[cap + i];
        v[0] = (t == 0.0) ? x1 / R.p[0] : 0.0;
        v[1] = (t == 0.0 && R.p[3] > x1 && x1 > R.p[2]) ? x2 / R.p[1] : 0.0;
    } else if (RED == SP_RED_ENERGY_COLLISION) {  // fields {v, rho, rho0}; params {m, c, rho0}
        const double* V = R.f[0];
        double m = R.p[0], c = R.p[1], rho0 = R.p[2];
        double vx = V[i], vy = V[cap + i], vz = V[2 * cap + i];
        double d = R.f[1][i] - R.f[2][i];
        v[0] = 0.5 * m * (vx * vx + vy * vy + vz * vz) + 0.5 * m * (c * c) * (d * d) / (rho0 * rho0);
    } else if (RED == SP_RED_SUM) {
        for (int c = 0; c < R.ncomp && c < 3; c++) v[c] = R.f[0][c * cap + i];
    } else if (RED == SP_RED_ENERGY_ISPH) {  // fields {x, v}; params {m, gx, gy, gz}
        const double *X = R.f[0], *V = R.f[1];
        double m = R.p[0];
        double vx = V[i], vy = V[cap + i], vz = V[2 * cap + i];
        v[0] = 0.5 * m * (vx * vx + vy * vy + vz * vz) - m * (R.p[1] * X[i] + R.p[2] * X[cap + i] + R.p[3] * X[2 * cap + i]);
    } else if (RED == SP_RED_ENERGY_ROD) {  // fields {v, A}; params {m, c_s, c_l}
        const double* V = R.f[0];
        const double m = R.p[0], c_s = R.p[1], c_l = R.p[2];
        const SpM2 A = sp_m2_load(R.f[1], cap, (int)i);
        const double d = fabs(sp_m2_det(A));
        double lam;
        const SpM2 G0 = sp_m2_dev(sp_m2_mul(sp_m2_trans(A), A), &lam);
        const double g33 = 1.0 - lam;
        const double n2 = sqrt(G0.a11 * G0.a11 + G0.a21 * G0.a21 + G0.a12 * G0.a12 + G0.a22 * G0.a22 + g33 * g33);
        const double vx = V[i], vy = V[cap + i], vz = V[2 * cap + i];
        v[0] = 0.5 * m * (vx * vx + vy * vy + vz * vz) + 0.25 * m * (c_s * c_s) * (n2 * n2) +
               m * (c_l * c_l) * (d - 1.0 - log(d));
    } else if (RED == SP_RED_MAX_SPEED) {  // fields {v}
        const double* V = R.f[0];
        const double vx = V[i], vy = V[cap + i], vz = V[2 * cap + i];
        v[0] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));  // norm, algebra.jl:58-60
    } else if (RED == SP_RED_FORCE_ON_TYPE) {  // fields {a, m, type}; params {type_sel}
        if (R.f[2][i] == R.p[0]) {
            const double* A = R.f[0];
            const double m = R.f[1][i];
            v[0] = m * A[i];
            v[1] = m * A[cap + i];
            v[2] = m * A[2 * cap + i];
        }
    }
}

template <bool IS_MAX>
__device__ __forceinline__ double red_op(double a, double b) {
    return IS_MAX ? fmax(a, b) : a + b;
}

template <bool IS_MAX>
__device__ __forceinline__ void block_reduce3(double v[3], double* out3) {
    __shared__ double sm[3][RED_B / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double x = v[c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x = red_op<IS_MAX>(x, __shfl_down_sync(0xffffffffu, x, d));
        if (lane == 0) sm[c][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double x = (lane < RED_B / 32) ? sm[c][lane] : 0.0;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) x = red_op<IS_MAX>(x, __shfl_down_sync(0xffffffffu, x, d));
            if (lane == 0) out3[c] = x;
        }
    }
}

template <int RED, bool IS_MAX>
__global__ void __launch_bounds__(RED_B) k_reduce(RedParams R, const int* __restrict__ alive, double* partial) {
    double acc[3] = {0.0, 0.0, 0.0};
    const long long n = *alive;  // the dead tail of culled particles takes no part
    for (long long i = blockIdx.x * (long long)RED_B + threadIdx.x; i < n; i += (long long)gridDim.x * RED_B) {
        double v[3] = {0.0, 0.0, 0.0};
        if (R.ghost && R.ghost[i] != 0.0) continue;
        red_map<RED>(R, i, v);
#pragma unroll
        for (int c = 0; c < 3; c++) acc[c] = red_op<IS_MAX>(acc[c], v[c]);
    }
    block_reduce3<IS_MAX>(acc, partial + 3 * blockIdx.x);
}
template <bool IS_MAX>
__global__ void __launch_bounds__(RED_B) k_reduce_final(const double* partial, int nblocks, double* out) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < nblocks; b += RED_B)
#pragma unroll
        for (int c = 0; c < 3; c++) acc[c] = red_op<IS_MAX>(acc[c], partial[3 * b + c]);
    block_reduce3<IS_MAX>(acc, out);
}

// Shared with sp_isph.cu: sum of a[i]*b[i] into a device scalar (deterministic two-stage).
__global__ void __launch_bounds__(RED_B) k_dot_partial(const double* a, const double* b, const double* ghost, long long n,
                                                       double* partial) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long i = blockIdx.x * (long long)RED_B + threadIdx.x; i < n; i += (long long)gridDim.x * RED_B)
        if (!ghost || ghost[i] == 0.0) acc[0] += a[i] * b[i];
    block_reduce3<false>(acc, partial + 3 * blockIdx.x);
}
int sp_dot_device(sp_system* s, const double* a, const double* b, long long n, double* partial, double* out3) {
    int nb = (int)((n + RED_B - 1) / RED_B);
    if (nb > RED_MAXBLOCKS) nb = RED_MAXBLOCKS;
    if (nb < 1) nb = 1;
    SP_LAUNCH(s, k_dot_partial, nb, RED_B, 0, a, b, sp_slab_ghost_mask(s), n, partial);
    SP_LAUNCH(s, k_reduce_final<false>, 1, RED_B, 0, partial, nb, out3);
    return SP_OK;
}

template <int RED, bool IS_MAX>
static int run_reduce(sp_system* s, const RedParams& R, int nout, double* out) {
    int rc = sp_ensure_stage(s, 3 * RED_MAXBLOCKS + 8);
    if (rc) return rc;
    int nb = (int)((s->n + RED_B - 1) / RED_B);
    if (nb > RED_MAXBLOCKS) nb = RED_MAXBLOCKS;
    if (nb < 1) nb = 1;
    double* partial = s->stage;
    double* res = s->stage + 3 * RED_MAXBLOCKS;
    SP_LAUNCH(s, (k_reduce<RED, IS_MAX>), nb, RED_B, 0, R, sp_alive(s), partial);
    SP_LAUNCH(s, (k_reduce_final<IS_MAX>), 1, RED_B, 0, partial, nb, res);
    if ((rc = sp_slab_allreduce_device(s, res, 3, IS_MAX ? 1 : 0))) return rc;
    double h[3];
    SP_CUDA(s, cudaMemcpyAsync(h, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    for (int c = 0; c < nout; c++) out[c] = h[c];
    return SP_OK;
}

// ------------------------------------------------------------------ point sums
struct PointSumParams {
    const double *x, *y, *z, *type, *f;
    const int* cell_start;
    double m, tsel;
    SpKC kc;
    int with_f;
};

template <class K>
__global__ void k_point_sum(SpGrid g, PointSumParams P, const double* pts, long long m_pts, double* out) {
    long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= m_pts) return;
    const double xi = pts[3 * k], yi = pts[3 * k + 1], zi = pts[3 * k + 2];
    const long long key = sp_find_key(g, xi, yi, zi);
    const long long L1 = g.lim[0], L12 = g.lim[0] * g.lim[1];
    const int nk = (g.dim == 2) ? 0 : 1;
    double acc = 0.0;
    // key_diff order, descending index inside a cell: the reference's summation order (core.jl:243-257)
    for (int di = -1; di <= 1; di++)
        for (int dj = -1; dj <= 1; dj++)
            for (int dk = -nk; dk <= nk; dk++) {
                const long long nkey = key + di + L1 * dj + L12 * dk;
                if (nkey < 1 || nkey > g.key_max) continue;
                for (int j = P.cell_start[nkey]; j < P.cell_start[nkey + 1]; j++) {
                    double dx = __dsub_rn(xi, P.x[j]), dy = __dsub_rn(yi, P.y[j]), dz = __dsub_rn(zi, P.z[j]);
                    double d2 = sp_d2(dx, dy, dz);
                    if (d2 > g.T2) continue;  // no self exclusion
                    double sel = (P.type[j] == P.tsel) ? 1.0 : 0.0;
                    double w = K::w(P.kc, sqrt(d2));
                    acc += P.with_f ? sel * P.m * P.f[j] * w : sel * P.m * w;
                }
            }
    out[k] = acc;
}

// ------------------------------------------------------------------ kernel evaluation
template <class K>
__global__ void k_kernel_eval(SpKC kc, int kfun, const double* r, double* out, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v;
    switch (kfun) {
        case SP_KFUN_W: v = K::w(kc, r[i]); break;
        case SP_KFUN_DW: v = K::D(kc, r[i]); break;
        case SP_KFUN_RDW: v = K::rD(kc, r[i]); break;
        default: v = K::DD(kc, r[i]); break;
    }
    out[i] = v;
}

extern "C" {

int32_t sp_reduce(sp_system* s, int32_t red, const int32_t* F, int32_t nf, const double* Pm, int32_t np, double* out) {
    if (!s || !out) return SP_ERR_INVALID;
    SP_NOT_WHILE_RECORDING(s);
    SP_CUDA(s, cudaSetDevice(s->device));
    RedParams R{};
    R.cap = s->cap;
    R.ghost = sp_slab_ghost_mask(s);
    auto bind = [&](int nexp, const int* nc, int npar) -> int {
        int rc = sp_check_fields(s, F, nf, nc, nexp);
        if (rc) return rc;
        if (np != npar || (npar && !Pm)) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters for this reduction");
        for (int i = 0; i < nexp; i++) R.f[i] = s->fields[F[i]].d;
        for (int i = 0; i < npar; i++) R.p[i] = Pm[i];
        return SP_OK;
    };
    int rc;
    switch (red) {
        case SP_RED_ENERGY_WCSPH: {
            const int nc[] = {3, 3, 1};
            if ((rc = bind(3, nc, 6))) return rc;
            return run_reduce<SP_RED_ENERGY_WCSPH, false>(s, R, 1, out);
        }
        case SP_RED_FRONT: {
            const int nc[] = {3, 1};
            if ((rc = bind(2, nc, 4))) return rc;
            return run_reduce<SP_RED_FRONT, true>(s, R, 2, out);
        }
        case SP_RED_ENERGY_COLLISION: {
            const int nc[] = {3, 1, 1};
            if ((rc = bind(3, nc, 3))) return rc;
            return run_reduce<SP_RED_ENERGY_COLLISION, false>(s, R, 1, out);
        }
        case SP_RED_SUM: {
            const int nc[] = {0};
            if ((rc = bind(1, nc, 0))) return rc;
            R.ncomp = s->fields[F[0]].ncomp;
            if (R.ncomp > 3) return sp_fail(s, SP_ERR_INVALID, "SP_RED_SUM supports 1 or 3 components");
            return run_reduce<SP_RED_SUM, false>(s, R, R.ncomp, out);
        }
        case SP_RED_ENERGY_ISPH: {
            const int nc[] = {3, 3};
            if ((rc = bind(2, nc, 4))) return rc;
            return run_reduce<SP_RED_ENERGY_ISPH, false>(s, R, 1, out);
        }
        case SP_RED_ENERGY_ROD: {
            const int nc[] = {3, 9};
            if ((rc = bind(2, nc, 3))) return rc;
            return run_reduce<SP_RED_ENERGY_ROD, false>(s, R, 1, out);
        }
        case SP_RED_MAX_SPEED: {
            const int nc[] = {3};
            if ((rc = bind(1, nc, 0))) return rc;
            return run_reduce<SP_RED_MAX_SPEED, true>(s, R, 1, out);
        }
        case SP_RED_FORCE_ON_TYPE: {
            const int nc[] = {3, 1, 1};
            if ((rc = bind(3, nc, 1))) return rc;
            return run_reduce<SP_RED_FORCE_ON_TYPE, false>(s, R, 3, out);
        }
    }
    return sp_fail(s, SP_ERR_INVALID, "unknown reduction id");
}

int32_t sp_sum_at_points(sp_system* s, int32_t sum_op, const int32_t* F, int32_t nf, const double* Pm, int32_t np,
                         const double* xyz, int64_t m_pts, double* out) {
    if (!s || !xyz || !out || m_pts < 0) return SP_ERR_INVALID;
    SP_NOT_WHILE_RECORDING(s);
    SP_CUDA(s, cudaSetDevice(s->device));
    if (!s->have_cells) return sp_fail(s, SP_ERR_STATE, "no cell list: call sp_create_cell_list first");
    int rc;
    PointSumParams P{};
    if (sum_op == SP_SUM_MASS_W) {
        const int nc[] = {3, 1};
        if ((rc = sp_check_fields(s, F, nf, nc, 2))) return rc;
        if (np != 4 || !Pm) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters");
        P.with_f = 0;
    } else if (sum_op == SP_SUM_MASS_F_W) {
        const int nc[] = {3, 1, 0};
        if ((rc = sp_check_fields(s, F, nf, nc, 3))) return rc;
        if (np != 5 || !Pm) return sp_fail(s, SP_ERR_INVALID, "wrong number of parameters");
        int comp = (int)Pm[4];
        if (comp < 0 || comp >= s->fields[F[2]].ncomp) return sp_fail(s, SP_ERR_INVALID, "bad component");
        P.f = s->fields[F[2]].d + (size_t)comp * s->cap;
        P.with_f = 1;
    } else
        return sp_fail(s, SP_ERR_INVALID, "unknown point-sum id");
    if (F[0] != 0) return sp_fail(s, SP_ERR_INVALID, "the first field must be x (field 0)");
    if (m_pts == 0) return SP_OK;
    if (!sp_make_kc((int)Pm[0], Pm[2], &P.kc)) return sp_fail(s, SP_ERR_INVALID, "unknown SPH kernel id");
    const double* X = s->fields[0].d;
    P.x = X;
    P.y = X + s->cap;
    P.z = X + 2 * s->cap;
    P.type = s->fields[F[1]].d;
    P.cell_start = s->cell_start;
    P.m = Pm[1];
    P.tsel = Pm[3];
    if ((rc = sp_ensure_stage(s, 4 * m_pts))) return rc;
    double* d_pts = s->stage;
    double* d_out = s->stage + 3 * m_pts;
    SP_CUDA(s, cudaMemcpyAsync(d_pts, xyz, (size_t)3 * m_pts * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    const int kernel = (int)Pm[0];
    if (kernel == SP_KERNEL_SPLINE23)
        SP_LAUNCH(s, k_point_sum<KSpline23>, sp_blocks(m_pts, 64), 64, 0, s->g, P, d_pts, (long long)m_pts, d_out);
    else if (kernel == SP_KERNEL_SPLINE24)
        SP_LAUNCH(s, k_point_sum<KSpline24>, sp_blocks(m_pts, 64), 64, 0, s->g, P, d_pts, (long long)m_pts, d_out);
    else
        SP_LAUNCH(s, k_point_sum<KWendland>, sp_blocks(m_pts, 64), 64, 0, s->g, P, d_pts, (long long)m_pts, d_out);
    SP_CUDA(s, cudaMemcpyAsync(out, d_out, (size_t)m_pts * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    SP_CUDA(s, cudaStreamSynchronize(s->stream));
    return SP_OK;
}

int32_t sp_kernel_eval(int32_t kernel, int32_t kfun, double h, const double* r, double* out, int64_t n, int32_t device) {
    if (!r || !out || n < 0) return SP_ERR_INVALID;
    SpKC kc;
    if (!sp_make_kc(kernel, h, &kc)) return sp_fail(nullptr, SP_ERR_INVALID, "unknown SPH kernel id");
    if (kfun < SP_KFUN_W || kfun > SP_KFUN_DDW) return sp_fail(nullptr, SP_ERR_INVALID, "unknown kernel function id");
    if (n == 0) return SP_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return sp_fail_cuda(nullptr, e, "cudaSetDevice", __FILE__, __LINE__);
    double *dr = nullptr, *dout = nullptr;
    if ((e = cudaMalloc(&dr, (size_t)n * sizeof(double))) != cudaSuccess)
        return sp_fail_cuda(nullptr, e, "cudaMalloc", __FILE__, __LINE__);
    if ((e = cudaMalloc(&dout, (size_t)n * sizeof(double))) != cudaSuccess) {
        cudaFree(dr);
        return sp_fail_cuda(nullptr, e, "cudaMalloc", __FILE__, __LINE__);
    }
    cudaMemcpy(dr, r, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
    const unsigned nb = sp_blocks(n, 256);
    if (kernel == SP_KERNEL_SPLINE23) k_kernel_eval<KSpline23><<<nb, 256>>>(kc, kfun, dr, dout, n);
    else if (kernel == SP_KERNEL_SPLINE24) k_kernel_eval<KSpline24><<<nb, 256>>>(kc, kfun, dr, dout, n);
    else k_kernel_eval<KWendland><<<nb, 256>>>(kc, kfun, dr, dout, n);
    e = cudaMemcpy(out, dout, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dr);
    cudaFree(dout);
    if (e != cudaSuccess) return sp_fail_cuda(nullptr, e, "kernel_eval", __FILE__, __LINE__);
    return SP_OK;
}

}  // extern "C"
