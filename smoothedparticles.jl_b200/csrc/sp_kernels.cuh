// sp_kernels.cuh — SPH kernel functions of src/kernels.jl as device functions.
//
// The reference marks all of them @fastmath, so only a tolerance is meaningful (the parity tests use
// 1e-13 relative against the oracle).  The h-dependent factors are folded on the host into SpKC once
// per call, and x = r/h is evaluated as r*(1/h): no division on the pair path except rDspline23's 1/x.
// Constants are the literal decimals of the source.
#pragma once
#include "sp_internal.cuh"

struct SpKC {
    double inv_h;
    double cw;   // coefficient of w   / h^dim
    double cD;   // coefficient of Dw  / h^(dim+1)
    double crD;  // coefficient of rDw / h^(dim+2)
    double cw2;  // spline23 outer-branch coefficient of w / h^2
};

static inline double sp_ipow(double h, int k) {
    double r = 1.0;
    for (int i = 0; i < k; i++) r *= h;
    return r;
}

static inline bool sp_make_kc(int kernel, double h, SpKC* kc) {
    kc->inv_h = 1.0 / h;
    kc->cw2 = 0.0;
    switch (kernel) {
        case SP_KERNEL_WENDLAND1:  // kernels.jl:206-228
            kc->cw = 1.5 / h;
            kc->cD = -30.0 / sp_ipow(h, 2);
            kc->crD = -30.0 / sp_ipow(h, 3);
            return true;
        case SP_KERNEL_WENDLAND2:  // kernels.jl:108-147
            kc->cw = 2.228169203286535 / sp_ipow(h, 2);
            kc->cD = -44.563384065730695 / sp_ipow(h, 3);
            kc->crD = -44.563384065730695 / sp_ipow(h, 4);
            return true;
        case SP_KERNEL_WENDLAND3:  // kernels.jl:156-204
            kc->cw = 3.3422538049298023 / sp_ipow(h, 3);
            kc->cD = -66.84507609859604 / sp_ipow(h, 4);
            kc->crD = -66.84507609859604 / sp_ipow(h, 5);
            return true;
        case SP_KERNEL_SPLINE23:  // kernels.jl:14-60
            kc->cw = 1.8189136353359467 / sp_ipow(h, 2);
            kc->cw2 = 3.6378272706718935 / sp_ipow(h, 2);
            kc->cD = -10.91348181201568 / sp_ipow(h, 3);
            kc->crD = -10.91348181201568 / sp_ipow(h, 4);
            return true;
        case SP_KERNEL_SPLINE24:  // kernels.jl:69-99
            kc->cw = 6.222175110452539 / sp_ipow(h, 2);
            kc->cD = -24.888700441810155 / sp_ipow(h, 3);
            kc->crD = -24.888700441810155 / sp_ipow(h, 4);
            return true;
    }
    return false;
}

#ifdef __CUDACC__
__device__ __forceinline__ double sp_pos(double x) { return x > 0.0 ? x : 0.0; }  // kernels.jl:3-5

// Wendland quintic, any dimension: only the folded coefficients differ.
struct KWendland {
    __device__ static __forceinline__ double w(const SpKC& k, double r) {
        double x = r * k.inv_h;
        if (x > 1.0) return 0.0;
        double t = 1.0 - x, t2 = t * t;
        return k.cw * (t2 * t2) * (1.0 + 4.0 * x);
    }
    __device__ static __forceinline__ double D(const SpKC& k, double r) {
        double x = r * k.inv_h;
        if (x > 1.0) return 0.0;
        double t = 1.0 - x;
        return k.cD * x * (t * t * t);
    }
    __device__ static __forceinline__ double rD(const SpKC& k, double r) {
        double x = r * k.inv_h;
        if (x > 1.0) return 0.0;
        double t = 1.0 - x;
        return k.crD * (t * t * t);
    }
    // DDwendland3, kernels.jl:197-204 (same 1/h^5 factor as rDwendland3)
    __device__ static __forceinline__ double DD(const SpKC& k, double r) {
        double x = r * k.inv_h;
        if (x > 1.0) return 0.0;
        double t = 1.0 - x;
        return k.crD * ((1.0 - 4.0 * x) * (t * t));
    }
};

struct KSpline23 {
    __device__ static __forceinline__ double w(const SpKC& k, double r) {
        double x = r * k.inv_h;
        if (x < 0.5) return k.cw * (1.0 - 6.0 * (x * x) + 6.0 * (x * x * x));
        if (x < 1.0) {
            double t = 1.0 - x;
            return k.cw2 * (t * t * t);
        }
        return 0.0;
    }
    __device__ static __forceinline__ double D(const SpKC& k, double r) {
        double x = r * k.inv_h;
        if (x < 0.5) return k.cD * (2.0 * x - 3.0 * (x * x));
        if (x < 1.0) {
            double t = 1.0 - x;
            return k.cD * (t * t);
        }
        return 0.0;
    }
    __device__ static __forceinline__ double rD(const SpKC& k, double r) {
        double x = r * k.inv_h;
        if (x < 0.5) return k.crD * (2.0 - 3.0 * x);
        if (x < 1.0) {
            double t = 1.0 - x;
            return k.crD * (t * t) / x;
        }
        return 0.0;
    }
    __device__ static __forceinline__ double DD(const SpKC&, double) { return 0.0; }
};

struct KSpline24 {
    __device__ static __forceinline__ double w(const SpKC& k, double r) {
        double x = r * k.inv_h;
        double a = sp_pos(1.0 - x), b = sp_pos(0.6 - x), c = sp_pos(0.2 - x);
        a *= a; b *= b; c *= c;
        return k.cw * (a * a - 5.0 * (b * b) + 10.0 * (c * c));
    }
    __device__ static __forceinline__ double D(const SpKC& k, double r) {
        double x = r * k.inv_h;
        double a = sp_pos(1.0 - x), b = sp_pos(0.6 - x), c = sp_pos(0.2 - x);
        return k.cD * (a * a * a - 5.0 * (b * b * b) + 10.0 * (c * c * c));
    }
    __device__ static __forceinline__ double rD(const SpKC& k, double r) {
        double x = r * k.inv_h;
        if (x > 0.2) {
            double a = sp_pos(1.0 - x), b = sp_pos(0.6 - x);
            return k.crD * (a * a * a - 5.0 * (b * b * b)) / x;
        }
        return k.crD * (1.2 - 6.0 * (x * x));
    }
    __device__ static __forceinline__ double DD(const SpKC&, double) { return 0.0; }
};
#endif
