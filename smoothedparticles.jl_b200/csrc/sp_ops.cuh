// sp_ops.cuh — the registered operators: device restatements of the per-pair / per-particle closures
// the reference's examples pass to apply!.  Each struct cites its source.
//
// Pair operator interface used by the sweep kernels (sp_sweep.cu):
//   Params            POD block: field plane pointers + folded constants
//   PS                per-p state loaded once (registers)
//   Acc               accumulators; init() loads the CURRENT value of the p-owned output so the
//                     accumulation order is the reference's:  ((old + t1) + t2) + ...
//   active(P,i)       type guard of the closure (skips the whole neighbour loop)
//   NQ, P.qp[NQ]      the q-side Float64 planes the pair body reads (beyond x); the sweep kernels hand
//                     pair() an accessor q(k) = value of plane k for the current neighbour, read either
//                     from HBM/L1 (reference-order kernel) or from the staged shared-memory tile
//   pair(P,p,q,dx,dy,dz,r,acc)    one accepted pair
//   self(P,p,acc)     the (p,p,0.0) term of apply!(...; self=true)  (core.jl:155-157)
//   store(P,i,p,acc)  write back p-owned outputs
#pragma once
#include "sp_kernels.cuh"

struct RV3 {
    const double *x, *y, *z;
};
struct WV3 {
    double *x, *y, *z;
};

// ---------------------------------------------------------------- WCSPH
// balance_of_mass!  collapse_dry.jl:112-115, collapse3d.jl:87-90, cavity_flow.jl:92-94 (two_nu = 0)
template <class K>
struct OpBalanceOfMass {
    static constexpr bool FUSED_BUILD = true;  // first pair sweep after a cell-list build: see k_nbr_build_sweep
    static constexpr int NQ = 4;  // vx, vy, vz, rho
    struct Params {
        const double* qp[NQ];
        double* Drho;
        double m, two_nu;
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz, rho;
    };
    struct Acc {
        double d;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        p.rho = P.qp[3][i];
        a.d = P.Drho[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double ker = P.m * K::rD(P.kc, r);
        double dvx = p.vx - q(0), dvy = p.vy - q(1), dvz = p.vz - q(2);
        a.d += ker * ((dx * dvx + dy * dvy + dz * dvz) + P.two_nu * (p.rho - q(3)));
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) { P.Drho[i] = a.d; }
};

// internal_force!  collapse_dry.jl:135-141 (3-D: collapse3d.jl:98-104 with the same formula)
template <class K>
struct OpInternalForce {
    static constexpr bool FUSED_BUILD = true;  // first pair sweep after a cell-list build: see k_nbr_build_sweep
    // The per-particle quotient P/rho^2 of both p and q is evaluated ONCE per particle by UPressureOverRho2
    // (same IEEE division as the closure's p.P/p.rho^2, just hoisted out of the pair loop).
    static constexpr int NQ = 4;  // vx, vy, vz, pr = P/rho^2
    struct Params {
        const double* qp[NQ];
        const double* type;
        WV3 Dv;
        double m, visc;  // visc = 2*mu/rho0^2
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz, pr;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params& P, int i) { return P.type[i] == 0.0; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        p.pr = P.qp[3][i];
        a.x = P.Dv.x[i]; a.y = P.Dv.y[i]; a.z = P.Dv.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double ker = P.m * K::rD(P.kc, r);
        double c = -ker * (p.pr + q(3));
        double b = ker * P.visc;
        a.x += c * dx; a.y += c * dy; a.z += c * dz;
        a.x += b * (p.vx - q(0)); a.y += b * (p.vy - q(1)); a.z += b * (p.vz - q(2));
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.Dv.x[i] = a.x; P.Dv.y[i] = a.y; P.Dv.z[i] = a.z;
    }
};

// ---- cross-operator cache for the WCSPH pair (balance_of_mass!, internal_force!) of one time step.
// internal_force! sums  -ker*(pr_p + pr_q)*x_pq + visc*ker*v_pq  over the same neighbours, with the same
// ker = m*rDw(h,r), x_pq and v_pq that balance_of_mass! has just evaluated (positions and velocities do not
// change in between, collapse3d.jl:136-150).  Splitting the sum,
//     Dv_p += -pr_p * A_p - B_p + visc * C_p,    A_p = sum ker*x_pq,  C_p = sum ker*v_pq,  B_p = sum ker*pr_q*x_pq,
// lets the mass sweep accumulate A and C on the side (6 FMAs per pair, no extra loads) and leaves the force sweep
// with B only: it gathers x and pr of q (4 planes instead of 7) — the gathers are what bounds the replay kernels.
// Same terms, different association: within the 1e-10 parity bar (measured ~1e-13); SP_FLAG_STRICT_ORDER and any
// change of x, v, kernel, m or h in between fall back to the plain operator.
template <class K>
struct OpBalanceOfMassAux {
    static constexpr bool FUSED_BUILD = true;  // first pair sweep after a cell-list build: see k_nbr_build_sweep
    static constexpr int NQ = 4;  // vx, vy, vz, rho
    struct Params {
        const double* qp[NQ];
        double* Drho;
        WV3 kx, kv;  // A and C
        double m, two_nu;
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz, rho;
    };
    struct Acc {
        double d, ax, ay, az, cx, cy, cz;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        p.rho = P.qp[3][i];
        a.d = P.Drho[i];
        a.ax = a.ay = a.az = a.cx = a.cy = a.cz = 0.0;
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double ker = P.m * K::rD(P.kc, r);
        double dvx = p.vx - q(0), dvy = p.vy - q(1), dvz = p.vz - q(2);
        a.d += ker * ((dx * dvx + dy * dvy + dz * dvz) + P.two_nu * (p.rho - q(3)));
        a.ax += ker * dx; a.ay += ker * dy; a.az += ker * dz;
        a.cx += ker * dvx; a.cy += ker * dvy; a.cz += ker * dvz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.Drho[i] = a.d;
        P.kx.x[i] = a.ax; P.kx.y[i] = a.ay; P.kx.z[i] = a.az;
        P.kv.x[i] = a.cx; P.kv.y[i] = a.cy; P.kv.z[i] = a.cz;
    }
};

template <class K>
struct OpInternalForceCached {
    static constexpr int NQ = 1;  // pr = P/rho^2
    struct Params {
        const double* qp[NQ];
        const double* type;
        WV3 Dv;
        RV3 kx, kv;
        double m, visc;  // visc = 2*mu/rho0^2
        SpKC kc;
    };
    struct PS {
        double pr;
    };
    struct Acc {
        double x, y, z;  // B
    };
    __device__ static __forceinline__ bool active(const Params& P, int i) { return P.type[i] == 0.0; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.pr = P.qp[0][i];
        a.x = a.y = a.z = 0.0;
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS&, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double c = (P.m * K::rD(P.kc, r)) * q(0);
        a.x += c * dx; a.y += c * dy; a.z += c * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS& p, const Acc& a) {
        P.Dv.x[i] += P.visc * P.kv.x[i] - (p.pr * P.kx.x[i] + a.x);
        P.Dv.y[i] += P.visc * P.kv.y[i] - (p.pr * P.kx.y[i] + a.y);
        P.Dv.z[i] += P.visc * P.kv.z[i] - (p.pr * P.kx.z[i] + a.z);
    }
};

// internal_force!  cavity_flow.jl:102-114 (rDwendland2; lid extrapolation; Monaghan viscosity)
template <class K>
struct OpInternalForceCavity {
    static constexpr bool FUSED_BUILD = true;  // first pair sweep after a cell-list build: see k_nbr_build_sweep
    static constexpr int NQ = 6;  // vx, vy, vz, pr = P/rho^2, rho, type
    struct Params {
        const double* qp[NQ];
        WV3 Dv;
        double m, Re, vlid, ylid, lid, tenth_h, eps;  // eps = 0.01*h^2
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz, pr, rho, ay;  // ay = 0.1*h + |p.x[2] - ylid|
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double yi, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        p.pr = P.qp[3][i];
        p.rho = P.qp[4][i];
        p.ay = P.tenth_h + fabs(yi - P.ylid);
        a.x = P.Dv.x[i]; a.y = P.Dv.y[i]; a.z = P.Dv.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double rDk = K::rD(P.kc, r);
        double vx = p.vx - q(0), vy = p.vy - q(1), vz = p.vz - q(2);
        if (q(5) == P.lid) {
            double s = fabs(dy) / p.ay;
            vx = s * (p.vx - P.vlid); vy = s * p.vy; vz = s * p.vz;
        }
        double rq = q(4);
        double c = -P.m * rDk * (p.pr + q(3));
        a.x += c * dx; a.y += c * dy; a.z += c * dz;
        double b = 8.0 / (P.Re * p.rho * rq) * P.m * rDk * (vx * dx + vy * dy + vz * dz) / (r * r + P.eps);
        a.x += b * dx; a.y += b * dy; a.z += b * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.Dv.x[i] = a.x; P.Dv.y[i] = a.y; P.Dv.z[i] = a.z;
    }
};

// ---------------------------------------------------------------- examples/static_container.jl
// balance_of_mass!  :102-104 — the density itself is integrated in the pair loop
template <class K>
struct OpScBalanceOfMass {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 3;  // vx, vy, vz
    struct Params {
        const double* qp[NQ];
        double* rho;
        double m, dt;
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz;
    };
    struct Acc {
        double d;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        a.d = P.rho[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double dvx = p.vx - q(0), dvy = p.vy - q(1), dvz = p.vz - q(2);
        a.d += P.dt * (dx * dvx + dy * dvy + dz * dvz) * P.m * K::rD(P.kc, r);
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) { P.rho[i] = a.d; }
};

// internal_force!  :106-114 with pressure(p) = c^2*(rho - rho0) (:68-70) hoisted to one evaluation per particle
template <class K>
struct OpScInternalForce {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 5;  // vx, vy, vz, pr = P(rho)/rho^2, rho
    struct Params {
        const double* qp[NQ];
        const double* type;
        WV3 a;
        double m, two_mu;
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz, pr, rho;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params& P, int i) { return P.type[i] == 0.0; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        p.pr = P.qp[3][i];
        p.rho = P.qp[4][i];
        a.x = P.a.x[i]; a.y = P.a.y[i]; a.z = P.a.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double ker = P.m * K::rD(P.kc, r);
        double c = -ker * (p.pr + q(3));
        a.x += c * dx; a.y += c * dy; a.z += c * dz;
        double b = ker * P.two_mu / (p.rho * q(4));
        a.x += b * (p.vx - q(0)); a.y += b * (p.vy - q(1)); a.z += b * (p.vz - q(2));
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.a.x[i] = a.x; P.a.y[i] = a.y; P.a.z[i] = a.z;
    }
};

// ---------------------------------------------------------------- examples/drop.jl
// find_n!  :76-78
template <class K>
struct OpFindNormal {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 0;
    struct Params {
        const double* qp[1];
        WV3 n;
        double coef;  // 2*vol*vol
        SpKC kc;
    };
    struct PS {};
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS&, Acc& a) {
        a.x = P.n.x[i]; a.y = P.n.y[i]; a.z = P.n.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS&, const Q&, double dx, double dy, double dz,
                                                double r, Acc& a) {
        double k = P.coef * K::rD(P.kc, r);
        a.x += k * dx; a.y += k * dy; a.z += k * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}  // (p,p,0): x_pp = 0 adds nothing
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.n.x[i] = a.x; P.n.y[i] = a.y; P.n.z[i] = a.z;
    }
};

// internal_force!  :101-113 (pressure with the constant rho0, viscosity, surface tension)
template <class K>
struct OpInternalForceTension {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 7;  // vx, vy, vz, P, nx, ny, nz
    struct Params {
        const double* qp[NQ];
        WV3 a;
        double m, mu, rho0sq, tens, s0;  // tens = 2*beta/rho0^2
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz, P, nx, ny, nz;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        p.P = P.qp[3][i];
        p.nx = P.qp[4][i]; p.ny = P.qp[5][i]; p.nz = P.qp[6][i];
        a.x = P.a.x[i]; a.y = P.a.y[i]; a.z = P.a.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double ker = P.m * K::rD(P.kc, r);
        double c = -ker * (p.P / P.rho0sq + q(3) / P.rho0sq);
        a.x += c * dx; a.y += c * dy; a.z += c * dz;
        double b = 2 * ker * P.mu / P.rho0sq;
        a.x += b * (p.vx - q(0)); a.y += b * (p.vy - q(1)); a.z += b * (p.vz - q(2));
        double nx = p.nx - q(4), ny = p.ny - q(5), nz = p.nz - q(6);
        double w = (P.m * K::DD(P.kc, r) - ker) * (dx * nx + dy * ny + dz * nz);
        double den = r * r + P.s0;
        a.x -= P.tens * (w * dx / den + ker * nx);
        a.y -= P.tens * (w * dy / den + ker * ny);
        a.z -= P.tens * (w * dz / den + ker * nz);
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.a.x[i] = a.x; P.a.y[i] = a.y; P.a.z[i] = a.z;
    }
};

// ---------------------------------------------------------------- examples/collapse_symplectic.jl, Kepler_vortex.jl
// s^4 as Julia evaluates an integer-literal power of a Float64 (Base.pow_body for n = 4: two squarings with their
// low parts carried along); at most 1 ulp from (s*s)*(s*s)
__device__ __forceinline__ double sp_julia_pow4(double x) {
    const double x2 = __dmul_rn(x, x), lo2 = __fma_rn(x, x, -x2);
    const double err = __dmul_rn(__dmul_rn(x2, 2.0), lo2);
    const double x4 = __dmul_rn(x2, x2);
    const double lo4 = __dadd_rn(__fma_rn(x2, x2, -x4), err);
    return (isfinite(x4) && isfinite(lo4)) ? __dadd_rn(x4, lo4) : x4;
}
// rev_add  examples/utils/FixPA.jl:28-30: 2^-30 * (Int64(round(x*2^30)) + Int64(round(y*2^30))), round half to even
__device__ __forceinline__ double sp_rev_add(double x, double y) {
    const long long a = __double2ll_rn(__dmul_rn(x, 1073741824.0)), b = __double2ll_rn(__dmul_rn(y, 1073741824.0));
    return __dmul_rn(__ll2double_rn(a + b), 1.0 / 1073741824.0);
}

// find_rho! / find_rho0!  collapse_symplectic.jl:98-108, Kepler_vortex.jl:139-149 (fluid-fluid pairs only; self=true)
template <class K>
struct OpDensitySumFluid {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 1;  // type
    struct Params {
        const double* qp[NQ];
        double* out;
        double m;
        SpKC kc;
    };
    struct PS {};
    struct Acc {
        double d;
    };
    __device__ static __forceinline__ bool active(const Params& P, int i) { return P.qp[0][i] == 0.0; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS&, Acc& a) {
        a.d = P.out[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS&, const Q& q, double, double, double, double r,
                                                Acc& a) {
        if (q(0) == 0.0) a.d += P.m * K::w(P.kc, r);
    }
    // (p, p, 0.0): only active (fluid) particles get here
    __device__ static __forceinline__ void self(const Params& P, const PS&, Acc& a) { a.d += P.m * K::w(P.kc, 0.0); }
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) { P.out[i] = a.d; }
};

// internal_force!  collapse_symplectic.jl:114-123, Kepler_vortex.jl:155-164: pressure between fluid particles,
// Lennard-Jones repulsion from wall particles closer than dr_wall.  pr is the hoisted per-particle quotient
// (P/rho^2 or P/rho0^2); it is 0/0 on wall particles of collapse_symplectic, where it is never used.
template <class K>
struct OpInternalForceLJ {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 2;  // pr, type
    struct Params {
        const double* qp[NQ];
        WV3 a;
        double m, wall, dr_wall, E_wall, eps;
        SpKC kc;
    };
    struct PS {
        double pr;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params& P, int i) { return P.qp[1][i] == 0.0; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.pr = P.qp[0][i];
        a.x = P.a.x[i]; a.y = P.a.y[i]; a.z = P.a.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        const double tq = q(1);
        if (tq == 0.0) {
            const double c = -(P.m * K::rD(P.kc, r)) * (p.pr + q(0));
            a.x += c * dx; a.y += c * dy; a.z += c * dz;
        } else if (tq == P.wall && r < P.dr_wall) {
            const double re = r + P.eps;
            const double s = P.dr_wall / re;
            const double c = -P.E_wall / (re * re) * (s * s - sp_julia_pow4(s));
            a.x += c * dx; a.y += c * dy; a.z += c * dz;
        }
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.a.x[i] = a.x; P.a.y[i] = a.y; P.a.z[i] = a.z;
    }
};

// sum(sys, LJ_potential, p) for every p  (core.jl:271-291 with collapse_symplectic.jl:146-153, Kepler_vortex.jl:186-193)
template <class K>
struct OpLJPotential {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 1;  // type
    struct Params {
        const double* qp[NQ];
        double* out;
        double coef, wall, dr_wall, eps;
        SpKC kc;
    };
    struct PS {};
    struct Acc {
        double d;
    };
    __device__ static __forceinline__ bool active(const Params& P, int i) { return P.qp[0][i] == 0.0; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS&, Acc& a) {
        a.d = P.out[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS&, const Q& q, double, double, double, double r,
                                                Acc& a) {
        if (q(0) == P.wall && r < P.dr_wall) {
            const double s = P.dr_wall / (r + P.eps);
            a.d += P.coef * (0.5 * (s * s) - 0.25 * sp_julia_pow4(s) - 0.25);
        }
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) { P.out[i] = a.d; }
};

// ---------------------------------------------------------------- examples/cylinder.jl (per-particle mass q.m)
// balance_of_mass!  :102-108
template <class K>
struct OpCylBalanceOfMass {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 6;  // vx, vy, vz, rho, m, type
    struct Params {
        const double* qp[NQ];
        double* Drho;
        double two_nu;
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz, rho, knu, tp;  // knu = two_nu/rho_p, the quotient the closure forms for every pair
    };
    struct Acc {
        double d;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        p.rho = P.qp[3][i];
        p.knu = P.two_nu / p.rho;
        p.tp = P.qp[5][i];
        a.d = P.Drho[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        const double ker = q(4) * K::rD(P.kc, r);
        const double dvx = p.vx - q(0), dvy = p.vy - q(1), dvz = p.vz - q(2);
        a.d += ker * (dx * dvx + dy * dvy + dz * dvz);
        if (p.tp == 0.0 && q(5) == 0.0) a.d += p.knu * (p.rho - q(3));
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) { P.Drho[i] = a.d; }
};

// internal_force!  :118-123 (pressure + Monaghan viscosity, every particle type)
template <class K>
struct OpCylInternalForce {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 6;  // vx, vy, vz, pr = P/rho^2, rho, m
    struct Params {
        const double* qp[NQ];
        WV3 a;
        double mu, eps2;  // eps2 = 0.01*h*h
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz, pr, rho;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        p.pr = P.qp[3][i];
        p.rho = P.qp[4][i];
        a.x = P.a.x[i]; a.y = P.a.y[i]; a.z = P.a.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        const double ker = q(5) * K::rD(P.kc, r);
        const double c = -ker * (p.pr + q(3));
        a.x += c * dx; a.y += c * dy; a.z += c * dz;
        const double dvx = p.vx - q(0), dvy = p.vy - q(1), dvz = p.vz - q(2);
        const double b = 8.0 * ker * P.mu / (p.rho * q(4)) * (dvx * dx + dvy * dy + dvz * dz) / (r * r + P.eps2);
        a.x += b * dx; a.y += b * dy; a.z += b * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.a.x[i] = a.x; P.a.y[i] = a.y; P.a.z[i] = a.z;
    }
};

// ---------------------------------------------------------------- examples/rod.jl (tensor-valued fields)
// In-plane block of a RealMatrix field (9 planes, Julia's column-major order: plane c = (i-1) + 3*(j-1)); rod.jl's own
// 2-D outer/det/inv/trans/dev (:44-85) populate nothing else.
struct SpM2 {
    double a11, a21, a12, a22;
};
__device__ __forceinline__ SpM2 sp_m2_load(const double* f, long long cap, int i) {
    return SpM2{f[i], f[cap + i], f[3 * cap + i], f[4 * cap + i]};
}
__device__ __forceinline__ void sp_m2_store(double* f, long long cap, int i, const SpM2& a) {
    f[i] = a.a11; f[cap + i] = a.a21; f[3 * cap + i] = a.a12; f[4 * cap + i] = a.a22;
    f[2 * cap + i] = 0.0; f[5 * cap + i] = 0.0; f[6 * cap + i] = 0.0; f[7 * cap + i] = 0.0; f[8 * cap + i] = 0.0;
}
__device__ __forceinline__ double sp_m2_det(const SpM2& a) { return a.a11 * a.a22 - a.a12 * a.a21; }
__device__ __forceinline__ SpM2 sp_m2_inv(const SpM2& a) {
    const double idet = 1.0 / sp_m2_det(a);
    return SpM2{idet * a.a22, -idet * a.a21, -idet * a.a12, idet * a.a11};
}
__device__ __forceinline__ SpM2 sp_m2_trans(const SpM2& a) { return SpM2{a.a11, a.a12, a.a21, a.a22}; }
__device__ __forceinline__ SpM2 sp_m2_mul(const SpM2& a, const SpM2& b) {
    return SpM2{a.a11 * b.a11 + a.a12 * b.a21, a.a21 * b.a11 + a.a22 * b.a21, a.a11 * b.a12 + a.a12 * b.a22,
                a.a21 * b.a12 + a.a22 * b.a22};
}
__device__ __forceinline__ SpM2 sp_m2_scale(double c, const SpM2& a) { return SpM2{c * a.a11, c * a.a21, c * a.a12, c * a.a22}; }
__device__ __forceinline__ SpM2 sp_m2_add(const SpM2& a, const SpM2& b) {
    return SpM2{a.a11 + b.a11, a.a21 + b.a21, a.a12 + b.a12, a.a22 + b.a22};
}
__device__ __forceinline__ SpM2 sp_m2_dev(const SpM2& g, double* lam_out) {
    const double lam = 1.0 / 3.0 * (g.a11 + g.a22 + 1.0);
    if (lam_out) *lam_out = lam;
    return SpM2{g.a11 - lam, g.a21, g.a12, g.a22 - lam};
}

// find_A!  rod.jl:128-134
template <class K>
struct OpRodFindA {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 2;  // X1, X2
    struct Params {
        const double* qp[NQ];
        double *A, *H;
        long long cap;
        SpKC kc;
    };
    struct PS {
        double X1, X2;
    };
    struct Acc {
        double a11, a21, a12, a22, h11, h21, h12, h22;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.X1 = P.qp[0][i]; p.X2 = P.qp[1][i];
        const SpM2 A = sp_m2_load(P.A, P.cap, i), H = sp_m2_load(P.H, P.cap, i);
        a.a11 = A.a11; a.a21 = A.a21; a.a12 = A.a12; a.a22 = A.a22;
        a.h11 = H.a11; a.h21 = H.a21; a.h12 = H.a12; a.h22 = H.a22;
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy, double,
                                                double r, Acc& a) {
        const double nk = -K::w(P.kc, r);
        const double X1 = p.X1 - q(0), X2 = p.X2 - q(1);
        a.a11 += nk * (X1 * dx); a.a21 += nk * (X2 * dx); a.a12 += nk * (X1 * dy); a.a22 += nk * (X2 * dy);
        a.h11 += nk * (dx * dx); a.h21 += nk * (dy * dx); a.h12 += nk * (dx * dy); a.h22 += nk * (dy * dy);
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        sp_m2_store(P.A, P.cap, i, SpM2{a.a11, a.a21, a.a12, a.a22});
        sp_m2_store(P.H, P.cap, i, SpM2{a.h11, a.h21, a.h12, a.h22});
    }
};

// find_f!  rod.jl:145-160
template <class K>
struct OpRodFindF {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 13;  // A11 A21 A12 A22 | B11 B21 B12 B22 | X1 X2 | vx vy vz
    struct Params {
        const double* qp[NQ];
        WV3 f;
        double two_m_vol, nu;
        SpKC kc;
    };
    struct PS {
        SpM2 A, B;
        double X1, X2, vx, vy, vz;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.A = SpM2{P.qp[0][i], P.qp[1][i], P.qp[2][i], P.qp[3][i]};
        p.B = SpM2{P.qp[4][i], P.qp[5][i], P.qp[6][i], P.qp[7][i]};
        p.X1 = P.qp[8][i]; p.X2 = P.qp[9][i];
        p.vx = P.qp[10][i]; p.vy = P.qp[11][i]; p.vz = P.qp[12][i];
        a.x = P.f.x[i]; a.y = P.f.y[i]; a.z = P.f.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy, double,
                                                double r, Acc& a) {
        const double ker = K::w(P.kc, r), rDker = K::rD(P.kc, r);
        const SpM2 Aq{q(0), q(1), q(2), q(3)}, Bq{q(4), q(5), q(6), q(7)};
        const double X1 = p.X1 - q(8), X2 = p.X2 - q(9);
        // -ker*(A'*(B*x_pq)) for p and for q
        double y1 = p.B.a11 * dx + p.B.a12 * dy, y2 = p.B.a21 * dx + p.B.a22 * dy;
        a.x += -ker * (p.A.a11 * y1 + p.A.a21 * y2);
        a.y += -ker * (p.A.a12 * y1 + p.A.a22 * y2);
        y1 = Bq.a11 * dx + Bq.a12 * dy; y2 = Bq.a21 * dx + Bq.a22 * dy;
        a.x += -ker * (Aq.a11 * y1 + Aq.a21 * y2);
        a.y += -ker * (Aq.a12 * y1 + Aq.a22 * y2);
        // "eta" correction: k_pq = +B_p'*(X_pq - A_p*x_pq), k_qp = -B_q'*(X_pq - A_q*x_pq)
        double w1 = X1 - (p.A.a11 * dx + p.A.a12 * dy), w2 = X2 - (p.A.a21 * dx + p.A.a22 * dy);
        const double kp1 = p.B.a11 * w1 + p.B.a21 * w2, kp2 = p.B.a12 * w1 + p.B.a22 * w2;
        w1 = X1 - (Aq.a11 * dx + Aq.a12 * dy); w2 = X2 - (Aq.a21 * dx + Aq.a22 * dy);
        const double kq1 = -(Bq.a11 * w1 + Bq.a21 * w2), kq2 = -(Bq.a12 * w1 + Bq.a22 * w2);
        const double dp = dx * kp1 + dy * kp2, dq = dx * kq1 + dy * kq2;
        a.x += rDker * dp * dx + ker * kp1;
        a.y += rDker * dp * dy + ker * kp2;
        a.x -= rDker * dq * dx + ker * kq1;
        a.y -= rDker * dq * dy + ker * kq2;
        // artificial viscosity
        const double visc = P.two_m_vol * rDker * P.nu;
        a.x += visc * (p.vx - q(10)); a.y += visc * (p.vy - q(11)); a.z += visc * (p.vz - q(12));
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.f.x[i] = a.x; P.f.y[i] = a.y; P.f.z[i] = a.z;
    }
};

// find_e!  rod.jl:185-188
template <class K>
struct OpRodFindE {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 2;  // X1, X2
    struct Params {
        const double* qp[NQ];
        const double* A;
        double* e;
        long long cap;
        SpKC kc;
    };
    struct PS {
        double X1, X2;
        SpM2 Ai;  // inv(A_p): the closure recomputes it for every pair, same value
    };
    struct Acc {
        double e;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.X1 = P.qp[0][i]; p.X2 = P.qp[1][i];
        p.Ai = sp_m2_inv(sp_m2_load(P.A, P.cap, i));
        a.e = P.e[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params&, const PS& p, const Q& q, double dx, double dy, double dz,
                                                double, Acc& a) {
        const double X1 = p.X1 - q(0), X2 = p.X2 - q(1);
        const double e1 = (p.Ai.a11 * X1 + p.Ai.a12 * X2) - dx, e2 = (p.Ai.a21 * X1 + p.Ai.a22 * X2) - dy, e3 = 0.0 - dz;
        a.e += e1 * e1 + e2 * e2 + e3 * e3;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) { P.e[i] = a.e; }
};

// ---------------------------------------------------------------- examples/SHTC/ldc.jl (full 3x3 RealMatrix fields)
// a[i + 3*j] = M[i+1, j+1] (Julia's column-major order; plane c of a 9-component field)
struct SpM3 {
    double a[9];
};
__device__ __forceinline__ SpM3 sp_m3_load(const double* f, long long cap, int i) {
    SpM3 m;
#pragma unroll
    for (int c = 0; c < 9; c++) m.a[c] = f[(size_t)c * cap + i];
    return m;
}
__device__ __forceinline__ void sp_m3_store(double* f, long long cap, int i, const SpM3& m) {
#pragma unroll
    for (int c = 0; c < 9; c++) f[(size_t)c * cap + i] = m.a[c];
}
__device__ __forceinline__ SpM3 sp_m3_mul(const SpM3& A, const SpM3& B) {
    SpM3 C;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) C.a[i + 3 * j] = A.a[i] * B.a[3 * j] + A.a[i + 3] * B.a[1 + 3 * j] + A.a[i + 6] * B.a[2 + 3 * j];
    return C;
}
__device__ __forceinline__ SpM3 sp_m3_tmul(const SpM3& A, const SpM3& B) {  // A'*B
    SpM3 C;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            C.a[i + 3 * j] = A.a[3 * i] * B.a[3 * j] + A.a[1 + 3 * i] * B.a[1 + 3 * j] + A.a[2 + 3 * i] * B.a[2 + 3 * j];
    return C;
}
__device__ __forceinline__ SpM3 sp_m3_scale(double c, const SpM3& A) {
    SpM3 C;
#pragma unroll
    for (int k = 0; k < 9; k++) C.a[k] = c * A.a[k];
    return C;
}
__device__ __forceinline__ SpM3 sp_m3_add(const SpM3& A, const SpM3& B) {
    SpM3 C;
#pragma unroll
    for (int k = 0; k < 9; k++) C.a[k] = A.a[k] + B.a[k];
    return C;
}
__device__ __forceinline__ SpM3 sp_m3_over(const SpM3& A, double n) {
    SpM3 C;
#pragma unroll
    for (int k = 0; k < 9; k++) C.a[k] = A.a[k] / n;
    return C;
}
__device__ __forceinline__ SpM3 sp_m3_dev(const SpM3& G) {  // ldc.jl:84-86
    const double lam = 1.0 / 3.0 * (G.a[0] + G.a[4] + G.a[8]);
    SpM3 C = G;
    C.a[0] = G.a[0] - lam; C.a[4] = G.a[4] - lam; C.a[8] = G.a[8] - lam;
    return C;
}

// update_v!  ldc.jl:123-127
template <class K>
struct OpShtcUpdateV {
    static constexpr bool LISTS_ONLY = true;  // see SpListsOnly, sp_sweep.cu
    static constexpr int NQ = 10;             // stress (9 planes), rho
    struct Params {
        const double* qp[NQ];
        const double* type;
        WV3 v;
        double dtm;
        SpKC kc;
    };
    struct PS {
        SpM3 Sp;  // stress_p/rho_p^2: the quotient the closure forms for every pair
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params& P, int i) { return P.type[i] == 0.0; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        const double r = P.qp[9][i], r2 = r * r;
#pragma unroll
        for (int c = 0; c < 9; c++) p.Sp.a[c] = P.qp[c][i] / r2;
        a.x = P.v.x[i]; a.y = P.v.y[i]; a.z = P.v.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        const double c = -P.dtm * K::rD(P.kc, r);
        const double rq = q(9), rq2 = rq * rq;
        double S[9];
#pragma unroll
        for (int k = 0; k < 9; k++) S[k] = c * (p.Sp.a[k] + q(k) / rq2);
        a.x += S[0] * dx + S[3] * dy + S[6] * dz;
        a.y += S[1] * dx + S[4] * dy + S[7] * dz;
        a.z += S[2] * dx + S[5] * dy + S[8] * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.v.x[i] = a.x; P.v.y[i] = a.y; P.v.z[i] = a.z;
    }
};

// update_rho!  ldc.jl:90-94
template <class K>
struct OpShtcUpdateRho {
    static constexpr bool LISTS_ONLY = true;
    static constexpr int NQ = 3;  // vx, vy, vz
    struct Params {
        const double* qp[NQ];
        const double* type;
        double* rho;
        double dtm;
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz;
    };
    struct Acc {
        double d;
    };
    __device__ static __forceinline__ bool active(const Params& P, int i) { return P.type[i] == 0.0; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        a.d = P.rho[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        a.d += P.dtm * K::rD(P.kc, r) * (dx * (p.vx - q(0)) + dy * (p.vy - q(1)) + dz * (p.vz - q(2)));
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) { P.rho[i] = a.d; }
};

// convect_A!  ldc.jl:96-100 — ORDER-DEPENDENT: every pair multiplies the A_p the previous pairs left, so the
// accumulator IS the running A_p and the sweep always runs in the reference's visiting order (strict kernel)
template <class K>
struct OpShtcConvectA {
    static constexpr bool LISTS_ONLY = true;
    static constexpr int NQ = 3;  // vx, vy, vz
    struct Params {
        const double* qp[NQ];
        const double *type, *rho;
        double* A;
        long long cap;
        double dtm, skip;
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz, s0;  // s0 = dtm/rho_p
    };
    struct Acc {
        double a[9];
    };
    __device__ static __forceinline__ bool active(const Params& P, int i) { return P.type[i] != P.skip; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        p.s0 = P.dtm / P.rho[i];
#pragma unroll
        for (int c = 0; c < 9; c++) a.a[c] = P.A[(size_t)c * P.cap + i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        const double s = p.s0 * K::rD(P.kc, r);
        const double v[3] = {p.vx - q(0), p.vy - q(1), p.vz - q(2)}, x[3] = {dx, dy, dz};
        SpM3 sA, M;
#pragma unroll
        for (int k = 0; k < 9; k++) sA.a[k] = s * a.a[k];
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int i = 0; i < 3; i++) M.a[i + 3 * j] = v[i] * x[j];  // v_pq*x_pq'
        const SpM3 D = sp_m3_mul(sA, M);
#pragma unroll
        for (int k = 0; k < 9; k++) a.a[k] += D.a[k];
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
#pragma unroll
        for (int c = 0; c < 9; c++) P.A[(size_t)c * P.cap + i] = a.a[c];
    }
};

// ---------------------------------------------------------------- examples/SHTC/beryllium.jl (SHTC solid, 2-D)
// the script's "structural" kernels wendland2h / rDwendland2h (:44-52): strict x < 1
__device__ __forceinline__ double sp_wendland2h(double h, double r) {
    const double x = r / h, u = 1.0 - x;
    return x < 1.0 ? 14.0 * (u * u * u) * (14.0 * (x * x) - 3.0 * x - 1.0) / (3.141592653589793 * (h * h)) : 0.0;
}
__device__ __forceinline__ double sp_rDwendland2h(double h, double r) {
    const double x = r / h, u = 1.0 - x, h2 = h * h;
    return x < 1.0 ? 140.0 * (u * u) * (4.0 - 7.0 * x) / (3.141592653589793 * (h2 * h2)) : 0.0;
}

// find_L!  :140-146 and find_J!  :153-158 share the T accumulation; WITH_L adds L, otherwise J and K
template <class K, bool WITH_L>
struct OpBeFindLJBase {
    static constexpr bool LISTS_ONLY = true;
    static constexpr int NQ = 3;  // m, vx, vy  (v only read when WITH_L)
    struct Params {
        const double* qp[NQ];
        double *T, *L, *J, *Kf;
        long long cap;
        double rho0, h;
        SpKC kc;
    };
    struct PS {
        double vx, vy, m;
    };
    struct Acc {
        double t11, t21, t12, t22, a, b, c, d;  // WITH_L: L11 L21 L12 L22; else: J, K, -, -
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        const SpM2 T = sp_m2_load(P.T, P.cap, i);
        a.t11 = T.a11; a.t21 = T.a21; a.t12 = T.a12; a.t22 = T.a22;
        p.m = P.qp[0][i];
        if (WITH_L) {
            p.vx = P.qp[1][i]; p.vy = P.qp[2][i];
            const SpM2 L = sp_m2_load(P.L, P.cap, i);
            a.a = L.a11; a.b = L.a21; a.c = L.a12; a.d = L.a22;
        } else {
            p.vx = p.vy = 0.0;
            a.a = P.J[i]; a.b = P.Kf[i]; a.c = a.d = 0.0;
        }
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy, double,
                                                double r, Acc& a) {
        const double mr = q(0) / P.rho0;
        const double ker = mr * K::rD(P.kc, r);
        a.t11 += ker * (dx * dx); a.t21 += ker * (dy * dx); a.t12 += ker * (dx * dy); a.t22 += ker * (dy * dy);
        if (WITH_L) {
            const double vx = p.vx - q(1), vy = p.vy - q(2);
            a.a += ker * (vx * dx); a.b += ker * (vy * dx); a.c += ker * (vx * dy); a.d += ker * (vy * dy);
        } else {
            a.a += mr * K::w(P.kc, r);
            a.b += mr * sp_wendland2h(P.h, r);
        }
    }
    // find_rho!(p, p, 0.0) of taco.jl:252 (self = true): x_pq = 0 adds nothing to T
    __device__ static __forceinline__ void self(const Params& P, const PS& p, Acc& a) {
        if (!WITH_L) {
            const double mr = p.m / P.rho0;
            a.a += mr * K::w(P.kc, 0.0);
            a.b += mr * sp_wendland2h(P.h, 0.0);
        }
    }
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        sp_m2_store(P.T, P.cap, i, SpM2{a.t11, a.t21, a.t12, a.t22});
        if (WITH_L) {
            sp_m2_store(P.L, P.cap, i, SpM2{a.a, a.b, a.c, a.d});
        } else {
            P.J[i] = a.a;
            P.Kf[i] = a.b;
        }
    }
};
template <class K>
struct OpBeFindL : OpBeFindLJBase<K, true> {};
template <class K>
struct OpBeFindJ : OpBeFindLJBase<K, false> {};

// find_f!  :166-175
template <class K>
struct OpBeFindF {
    static constexpr bool LISTS_ONLY = true;
    static constexpr int NQ = 6;  // m, K, T11, T21, T12, T22
    struct Params {
        const double* qp[NQ];
        WV3 f;
        double rho0, cp2, h;
        SpKC kc;
    };
    struct PS {
        double m, Kf;
        SpM2 T;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.m = P.qp[0][i]; p.Kf = P.qp[1][i];
        p.T = SpM2{P.qp[2][i], P.qp[3][i], P.qp[4][i], P.qp[5][i]};
        a.x = P.f.x[i]; a.y = P.f.y[i]; a.z = P.f.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        const double mr = q(0) / P.rho0;
        const double ker = mr * K::rD(P.kc, r), kerh = mr * sp_rDwendland2h(P.h, r);
        const double c = -p.m * ker;
        a.x += c * (p.T.a11 * dx + p.T.a12 * dy); a.y += c * (p.T.a21 * dx + p.T.a22 * dy);
        a.x += c * (q(2) * dx + q(4) * dy); a.y += c * (q(3) * dx + q(5) * dy);
        const double g = -p.m * kerh * P.cp2 * (p.Kf + q(1));
        a.x += g * dx; a.y += g * dy; a.z += g * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.f.x[i] = a.x; P.f.y[i] = a.y; P.f.z[i] = a.z;
    }
};

// ---------------------------------------------------------------- examples/SHTC/twist3d.jl (SHTC solid, 3-D)
__device__ __forceinline__ double sp_wendland3h(double h, double r) {  // :43-46, strict x < 1
    const double x = r / h, u = 1.0 - x;
    return x < 1.0 ? 21.0 * (u * u * u) * (14.0 * (x * x) - 3.0 * x - 1.0) / (3.141592653589793 * (h * h * h)) : 0.0;
}
__device__ __forceinline__ double sp_rDwendland3h(double h, double r) {  // :48-51
    const double x = r / h, u = 1.0 - x, h2 = h * h;
    return x < 1.0 ? 210.0 * (u * u) * (4.0 - 7.0 * x) / (3.141592653589793 * (h2 * h2 * h)) : 0.0;
}
__device__ __forceinline__ SpM3 sp_m3_inv(const SpM3& A) {  // adjugate / determinant
    const double* a = A.a;
    const double det = a[0] * a[4] * a[8] + a[1] * a[5] * a[6] + a[2] * a[3] * a[7] - a[6] * a[4] * a[2] - a[7] * a[5] * a[0] -
                       a[8] * a[3] * a[1];
    const double id = 1.0 / det;
    SpM3 C;
    C.a[0] = id * (a[4] * a[8] - a[7] * a[5]);
    C.a[1] = id * (a[7] * a[2] - a[1] * a[8]);
    C.a[2] = id * (a[1] * a[5] - a[4] * a[2]);
    C.a[3] = id * (a[6] * a[5] - a[3] * a[8]);
    C.a[4] = id * (a[0] * a[8] - a[6] * a[2]);
    C.a[5] = id * (a[3] * a[2] - a[0] * a[5]);
    C.a[6] = id * (a[3] * a[7] - a[6] * a[4]);
    C.a[7] = id * (a[6] * a[1] - a[0] * a[7]);
    C.a[8] = id * (a[0] * a[4] - a[3] * a[1]);
    return C;
}

// find_L!  :135-141 and find_J!  :148-153
template <class K, bool WITH_L>
struct OpTwFindLJBase {
    static constexpr bool LISTS_ONLY = true;
    static constexpr int NQ = 4;  // m, vx, vy, vz  (v only read when WITH_L)
    struct Params {
        const double* qp[NQ];
        double *T, *L, *J, *Kf;
        long long cap;
        double rho0, h;
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz;
    };
    struct Acc {
        double t[9], l[9];  // WITH_L: L; else l[0] = J, l[1] = K
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
#pragma unroll
        for (int c = 0; c < 9; c++) {
            a.t[c] = P.T[(size_t)c * P.cap + i];
            a.l[c] = WITH_L ? P.L[(size_t)c * P.cap + i] : 0.0;
        }
        if (WITH_L) {
            p.vx = P.qp[1][i]; p.vy = P.qp[2][i]; p.vz = P.qp[3][i];
        } else {
            p.vx = p.vy = p.vz = 0.0;
            a.l[0] = P.J[i];
            a.l[1] = P.Kf[i];
        }
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        const double mr = q(0) / P.rho0;
        const double ker = mr * K::rD(P.kc, r);
        const double x[3] = {dx, dy, dz};
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int i = 0; i < 3; i++) a.t[i + 3 * j] += ker * (x[i] * x[j]);
        if (WITH_L) {
            const double v[3] = {p.vx - q(1), p.vy - q(2), p.vz - q(3)};
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int i = 0; i < 3; i++) a.l[i + 3 * j] += ker * (v[i] * x[j]);
        } else {
            a.l[0] += mr * K::w(P.kc, r);
            a.l[1] += mr * sp_wendland3h(P.h, r);
        }
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
#pragma unroll
        for (int c = 0; c < 9; c++) {
            P.T[(size_t)c * P.cap + i] = a.t[c];
            if (WITH_L) P.L[(size_t)c * P.cap + i] = a.l[c];
        }
        if (!WITH_L) {
            P.J[i] = a.l[0];
            P.Kf[i] = a.l[1];
        }
    }
};
template <class K>
struct OpTwFindL : OpTwFindLJBase<K, true> {};
template <class K>
struct OpTwFindJ : OpTwFindLJBase<K, false> {};

// find_f!  :163-172
template <class K>
struct OpTwFindF {
    static constexpr bool LISTS_ONLY = true;
    static constexpr int NQ = 11;  // m, K, T (9 planes)
    struct Params {
        const double* qp[NQ];
        WV3 f;
        double rho0, cp2, h;
        SpKC kc;
    };
    struct PS {
        double m, Kf;
        double T[9];
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.m = P.qp[0][i]; p.Kf = P.qp[1][i];
#pragma unroll
        for (int c = 0; c < 9; c++) p.T[c] = P.qp[2 + c][i];
        a.x = P.f.x[i]; a.y = P.f.y[i]; a.z = P.f.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        const double mr = q(0) / P.rho0;
        const double ker = mr * K::rD(P.kc, r), kerh = mr * sp_rDwendland3h(P.h, r);
        const double c = p.m * ker;
        a.x += c * (p.T[0] * dx + p.T[3] * dy + p.T[6] * dz);
        a.y += c * (p.T[1] * dx + p.T[4] * dy + p.T[7] * dz);
        a.z += c * (p.T[2] * dx + p.T[5] * dy + p.T[8] * dz);
        a.x += c * (q(2) * dx + q(5) * dy + q(8) * dz);
        a.y += c * (q(3) * dx + q(6) * dy + q(9) * dz);
        a.z += c * (q(4) * dx + q(7) * dy + q(10) * dz);
        const double g = -p.m * kerh * P.cp2 * (p.Kf + q(1));
        a.x += g * dx; a.y += g * dy; a.z += g * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.f.x[i] = a.x; P.f.y[i] = a.y; P.f.z[i] = a.z;
    }
};

// ---------------------------------------------------------------- examples/SHTC/taco.jl
// find_f!  :154-162
template <class K>
struct OpTaFindF {
    static constexpr bool LISTS_ONLY = true;
    static constexpr int NQ = 11;  // m, lambda, T (9 planes)
    struct Params {
        const double* qp[NQ];
        WV3 f;
        double cpr2, h;
        SpKC kc;
    };
    struct PS {
        double m, lam;
        double T[9];
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.m = P.qp[0][i]; p.lam = P.qp[1][i];
#pragma unroll
        for (int c = 0; c < 9; c++) p.T[c] = P.qp[2 + c][i];
        a.x = P.f.x[i]; a.y = P.f.y[i]; a.z = P.f.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        const double mq = q(0);
        const double ker = mq * K::rD(P.kc, r), kerh = mq * sp_rDwendland2h(P.h, r);
        const double c = p.m * ker;
        double S[9];
#pragma unroll
        for (int k = 0; k < 9; k++) S[k] = c * (p.T[k] + q(2 + k));
        a.x += S[0] * dx + S[3] * dy + S[6] * dz;
        a.y += S[1] * dx + S[4] * dy + S[7] * dz;
        a.z += S[2] * dx + S[5] * dy + S[8] * dz;
        const double g = -p.m * kerh * P.cpr2 * (p.lam + q(1));
        a.x += g * dx; a.y += g * dy; a.z += g * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.f.x[i] = a.x; P.f.y[i] = a.y; P.f.z[i] = a.z;
    }
};

// ---------------------------------------------------------------- tests/test_collision_2d.jl
// find_rho! / find_rho0!  :63-69, used with self=true
template <class K>
struct OpDensitySum {
    static constexpr int NQ = 0;
    struct Params {
        const double* qp[1];
        double* out;
        double m;
        SpKC kc;
    };
    struct PS {};
    struct Acc {
        double d;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS&, Acc& a) {
        a.d = P.out[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS&, const Q&, double, double, double, double r,
                                                Acc& a) {
        a.d += P.m * K::w(P.kc, r);
    }
    __device__ static __forceinline__ void self(const Params& P, const PS&, Acc& a) { a.d += P.m * K::w(P.kc, 0.0); }
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) { P.out[i] = a.d; }
};

// internal_force!  test_collision_2d.jl:75-78
template <class K>
struct OpInternalForceSym {
    static constexpr int NQ = 1;  // P
    struct Params {
        const double* qp[NQ];
        WV3 a;
        double m, inv_rho0sq;
        SpKC kc;
    };
    struct PS {
        double pr;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.pr = P.qp[0][i] * P.inv_rho0sq;
        a.x = P.a.x[i]; a.y = P.a.y[i]; a.z = P.a.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double ker = P.m * K::rD(P.kc, r);
        double c = -ker * (p.pr + q(0) * P.inv_rho0sq);
        a.x += c * dx; a.y += c * dy; a.z += c * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.a.x[i] = a.x; P.a.y[i] = a.y; P.a.z[i] = a.z;
    }
};

// ---------------------------------------------------------------- ISPH, examples/collapse_dry_implicit.jl
// viscous_force!  :128-130
template <class K>
struct OpIsphViscous {
    static constexpr bool FUSED_BUILD = true;  // first pair sweep after a cell-list build: see k_nbr_build_sweep
    static constexpr int NQ = 3;  // vx, vy, vz
    struct Params {
        const double* qp[NQ];
        WV3 Dv;
        double coef;  // 2*m*mu/rho^2
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        a.x = P.Dv.x[i]; a.y = P.Dv.y[i]; a.z = P.Dv.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double, double, double,
                                                double r, Acc& a) {
        double c = P.coef * K::rD(P.kc, r);
        a.x += c * (p.vx - q(0)); a.y += c * (p.vy - q(1)); a.z += c * (p.vz - q(2));
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.Dv.x[i] = a.x; P.Dv.y[i] = a.y; P.Dv.z[i] = a.z;
    }
};

// div_L_lambda!  :147-152
template <class K>
struct OpIsphDivLLambda {
    static constexpr int NQ = 3;  // vx, vy, vz
    struct Params {
        const double* qp[NQ];
        double *div, *L, *lambda;
        double m, m_over_rho, inv_dim;
        SpKC kc;
    };
    struct PS {
        double vx, vy, vz;
    };
    struct Acc {
        double div, L, lam;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.vx = P.qp[0][i]; p.vy = P.qp[1][i]; p.vz = P.qp[2][i];
        a.div = P.div[i]; a.L = P.L[i]; a.lam = P.lambda[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double rDk = K::rD(P.kc, r);
        double dvx = p.vx - q(0), dvy = p.vy - q(1), dvz = p.vz - q(2);
        a.div += -(dx * dvx + dy * dvy + dz * dvz) * P.m * rDk;
        a.L += -2.0 * P.m_over_rho * rDk;
        a.lam += P.m_over_rho * rDk * (r * r) * P.inv_dim;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.div[i] = a.div; P.L[i] = a.L; P.lambda[i] = a.lam;
    }
};

// internal_force!  :132-134
template <class K>
struct OpIsphInternalForce {
    static constexpr int NQ = 1;  // P
    struct Params {
        const double* qp[NQ];
        WV3 Dv;
        double coef;  // m/rho^2
        SpKC kc;
    };
    struct PS {
        double P;
    };
    struct Acc {
        double x, y, z;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        p.P = P.qp[0][i];
        a.x = P.Dv.x[i]; a.y = P.Dv.y[i]; a.z = P.Dv.z[i];
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS& p, const Q& q, double dx, double dy,
                                                double dz, double r, Acc& a) {
        double c = P.coef * K::rD(P.kc, r) * (p.P + q(0));
        a.x -= c * dx; a.y -= c * dy; a.z -= c * dz;
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS&, const Acc& a) {
        P.Dv.x[i] = a.x; P.Dv.y[i] = a.y; P.Dv.z[i] = a.z;
    }
};

// Matrix-free (A p)_i of projection_matrix  :154-163 with assemble_matrix core.jl:196-225
//   y_i = A_ii p_i + sum_{j != i} (2 h^2 m/rho) rDk(r_ij) p_j,   A_ii = h^2 L_i + [type_i==0] C_free max(lambda_i,0)
template <class K>
struct OpPoissonApply {
    static constexpr int NQ = 1;  // p_in
    struct Params {
        const double* qp[NQ];
        const double *L, *lambda, *type;
        double* y;
        double off_coef;  // 2*h^2*m/rho
        double h2, C_free;
        SpKC kc;
    };
    struct PS {
        double diag;
    };
    struct Acc {
        double s;
    };
    __device__ static __forceinline__ bool active(const Params&, int) { return true; }
    __device__ static __forceinline__ void load(const Params& P, int i, double, double, double, PS& p, Acc& a) {
        double Aii = P.h2 * P.L[i];
        if (P.type[i] == 0.0) Aii += P.C_free * fmax(P.lambda[i], 0.0);
        p.diag = Aii * P.qp[0][i];
        a.s = 0.0;
    }
    template <class Q>
    __device__ static __forceinline__ void pair(const Params& P, const PS&, const Q& q, double, double, double,
                                                double r, Acc& a) {
        a.s += P.off_coef * K::rD(P.kc, r) * q(0);
    }
    __device__ static __forceinline__ void self(const Params&, const PS&, Acc&) {}
    __device__ static __forceinline__ void store(const Params& P, int i, const PS& p, const Acc& a) {
        P.y[i] = a.s + p.diag;
    }
};

// ---------------------------------------------------------------- unary operators
// pr = P/rho^2, the per-particle quotient of internal_force! (collapse_dry.jl:138, cavity_flow.jl:112)
struct UPressureOverRho2 {
    struct Params {
        const double *P, *rho;
        double* pr;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        double rho = P.rho[i];
        P.pr[i] = P.P[i] / (rho * rho);
    }
};
// find_pressure!  collapse_dry.jl:123-127, cavity_flow.jl:96-100
struct UFindPressure {
    struct Params {
        double *rho, *Drho, *P;
        double dt, c2, rho0, P0;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        double rho = P.rho[i] + P.Drho[i] * P.dt;
        P.rho[i] = rho;
        P.Drho[i] = 0.0;
        double pr = P.c2 * (rho - P.rho0);
        P.P[i] = (P.P0 != 0.0) ? P.P0 + pr : pr;
    }
};
// move!  collapse_dry.jl:148-153
struct UMove {
    struct Params {
        WV3 x;
        RV3 v;
        WV3 Dv;
        const double* type;
        double dtm;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        P.Dv.x[i] = 0.0; P.Dv.y[i] = 0.0; P.Dv.z[i] = 0.0;
        if (P.type[i] == 0.0) {
            P.x.x[i] += P.dtm * P.v.x[i]; P.x.y[i] += P.dtm * P.v.y[i]; P.x.z[i] += P.dtm * P.v.z[i];
        }
    }
};
// accelerate!  collapse_dry.jl:155-159
struct UAccelerate {
    struct Params {
        WV3 v;
        RV3 Dv;
        const double* type;
        double hdt, gx, gy, gz;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            P.v.x[i] += P.hdt * (P.Dv.x[i] + P.gx);
            P.v.y[i] += P.hdt * (P.Dv.y[i] + P.gy);
            P.v.z[i] += P.hdt * (P.Dv.z[i] + P.gz);
        }
    }
};
// ---- fused unary passes used by the step programs (sp_program.cu): the same statements in the same order as the
// separate operators (bit-identical results), one trip through HBM instead of two or three.
// accelerate!; accelerate!; move!  — the end of one collapse3d.jl step and the start of the next (:136-150)
struct UKickKickMove {
    struct Params {
        WV3 v, Dv, x;
        const double* type;
        double hdt, gx, gy, gz, dtm;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            const double ax = P.Dv.x[i] + P.gx, ay = P.Dv.y[i] + P.gy, az = P.Dv.z[i] + P.gz;
            double vx = P.v.x[i], vy = P.v.y[i], vz = P.v.z[i];
            vx += P.hdt * ax; vy += P.hdt * ay; vz += P.hdt * az;  // accelerate!
            vx += P.hdt * ax; vy += P.hdt * ay; vz += P.hdt * az;  // accelerate!
            P.v.x[i] = vx; P.v.y[i] = vy; P.v.z[i] = vz;
            P.x.x[i] += P.dtm * vx; P.x.y[i] += P.dtm * vy; P.x.z[i] += P.dtm * vz;  // move!
        }
        P.Dv.x[i] = 0.0; P.Dv.y[i] = 0.0; P.Dv.z[i] = 0.0;
    }
};
// find_pressure! followed by the hoisted P/rho^2 of internal_force! (UPressureOverRho2)
struct UFindPressurePr {
    struct Params {
        double *rho, *Drho, *P, *pr;
        double dt, c2, rho0, P0;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        double rho = P.rho[i] + P.Drho[i] * P.dt;
        P.rho[i] = rho;
        P.Drho[i] = 0.0;
        double pr = P.c2 * (rho - P.rho0);
        double pp = (P.P0 != 0.0) ? P.P0 + pr : pr;
        P.P[i] = pp;
        P.pr[i] = pp / (rho * rho);
    }
};
// pr = c2*(rho - rho0)/rho^2: the per-particle part of static_container.jl's internal_force! (:68-70, :110)
struct UEosPressureOverRho2 {
    struct Params {
        const double* rho;
        double* pr;
        double c2, rho0;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        double rho = P.rho[i];
        P.pr[i] = (P.c2 * (rho - P.rho0)) / (rho * rho);
    }
};
// move!  static_container.jl:116-119: every particle moves (walls have v = 0)
struct UMoveAll {
    struct Params {
        WV3 x;
        RV3 v;
        WV3 a;
        double dtm;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        P.x.x[i] += P.dtm * P.v.x[i]; P.x.y[i] += P.dtm * P.v.y[i]; P.x.z[i] += P.dtm * P.v.z[i];
        P.a.x[i] = 0.0; P.a.y[i] = 0.0; P.a.z[i] = 0.0;
    }
};
// normalize_n!  drop.jl:84-87
struct UNormalize {
    struct Params {
        WV3 n;
        double s0;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const double a = P.n.x[i], b = P.n.y[i], c = P.n.z[i];
        const double s = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c))) + P.s0;
        P.n.x[i] = a / s; P.n.y[i] = b / s; P.n.z[i] = c / s;
    }
};
// pr = P/rho0^2 with the constant rho0: the per-particle quotient of Kepler_vortex.jl:158
struct UPressureOverConst {
    struct Params {
        const double* P;
        double* pr;
        double rho0sq;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) { P.pr[i] = P.P[i] / P.rho0sq; }
};
// move!  collapse_symplectic.jl:134-138, Kepler_vortex.jl:174-178
struct UMoveRev {
    struct Params {
        WV3 x;
        RV3 v;
        const double* type;
        double dt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            P.x.x[i] = sp_rev_add(P.x.x[i], __dmul_rn(P.dt, P.v.x[i]));
            P.x.y[i] = sp_rev_add(P.x.y[i], __dmul_rn(P.dt, P.v.y[i]));
            P.x.z[i] = sp_rev_add(P.x.z[i], __dmul_rn(P.dt, P.v.z[i]));
        }
    }
};
// accelerate!  collapse_symplectic.jl:140-144
struct UAccelerateRev {
    struct Params {
        WV3 v;
        RV3 a;
        const double* type;
        double hdt, gx, gy, gz;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            P.v.x[i] = sp_rev_add(P.v.x[i], __dmul_rn(P.hdt, __dadd_rn(P.a.x[i], P.gx)));
            P.v.y[i] = sp_rev_add(P.v.y[i], __dmul_rn(P.hdt, __dadd_rn(P.a.y[i], P.gy)));
            P.v.z[i] = sp_rev_add(P.v.z[i], __dmul_rn(P.hdt, __dadd_rn(P.a.z[i], P.gz)));
        }
    }
};
// accelerate!  Kepler_vortex.jl:180-184 (central gravity -GM x/|x|^3 added reversibly to the SPH acceleration)
struct UAccelerateRevCentral {
    struct Params {
        RV3 x;
        WV3 v;
        RV3 a;
        const double* type;
        double hdt, GM;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            const double x = P.x.x[i], y = P.x.y[i], z = P.x.z[i];
            const double n = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
            const double k = -P.GM / __dmul_rn(__dmul_rn(n, n), n);
            P.v.x[i] = sp_rev_add(P.v.x[i], __dmul_rn(P.hdt, sp_rev_add(P.a.x[i], __dmul_rn(k, x))));
            P.v.y[i] = sp_rev_add(P.v.y[i], __dmul_rn(P.hdt, sp_rev_add(P.a.y[i], __dmul_rn(k, y))));
            P.v.z[i] = sp_rev_add(P.v.z[i], __dmul_rn(P.hdt, sp_rev_add(P.a.z[i], __dmul_rn(k, z))));
        }
    }
};
// find_pressure!  cylinder.jl:110-116 (the density of the inflow buffer upstream of x1_min is frozen)
struct UCylFindPressure {
    struct Params {
        const double* x;
        double *rho, *Drho, *P;
        double dt, c2, rho0, x1_min;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        double rho = P.rho[i];
        if (P.x[i] >= P.x1_min) {
            rho += P.Drho[i] * P.dt;
            P.rho[i] = rho;
        }
        P.Drho[i] = 0.0;
        P.P[i] = P.c2 * (rho - P.rho0);
    }
};
// move!  cylinder.jl:125-130
struct UMoveTypes {
    struct Params {
        WV3 x;
        RV3 v;
        WV3 a;
        const double* type;
        double dt, ta, tb;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        P.a.x[i] = 0.0; P.a.y[i] = 0.0; P.a.z[i] = 0.0;
        const double t = P.type[i];
        if (t == P.ta || t == P.tb) {
            P.x.x[i] += P.dt * P.v.x[i]; P.x.y[i] += P.dt * P.v.y[i]; P.x.z[i] += P.dt * P.v.z[i];
        }
    }
};
// accelerate! with the artificial attraction towards the cylinder  cylinder.jl:132-143
struct UCylAccelerate {
    struct Params {
        RV3 x;
        WV3 v;
        RV3 a;
        const double* type;
        double hdt, cyl1, coef;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            const double fx = P.cyl1 - P.x.x[i], fy = -P.x.y[i], x2 = P.x.y[i];
            const double absf2 = __dadd_rn(__dmul_rn(fx, fx), __dmul_rn(x2, x2));
            P.v.x[i] += P.hdt * (P.a.x[i] + P.coef * fx / absf2);
            P.v.y[i] += P.hdt * (P.a.y[i] + P.coef * fy / absf2);
            P.v.z[i] += P.hdt * (P.a.z[i] + P.coef * 0.0 / absf2);
        }
    }
};
// set_inflow_speed!  cylinder.jl:91-97
struct USetInflowSpeed {
    struct Params {
        RV3 x;
        WV3 v;
        const double* type;
        double inflow, s, U_max, chan_w;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == P.inflow) {
            const double q = 2.0 * P.x.y[i] / P.chan_w;
            const double v1 = __dmul_rn(__dmul_rn(P.s, P.U_max), __dsub_rn(1.0, __dmul_rn(q, q)));
            P.v.x[i] = v1; P.v.y[i] = v1 * 0.0; P.v.z[i] = v1 * 0.0;
        }
    }
};
// find_B!  rod.jl:136-143
struct URodFindB {
    struct Params {
        double *A, *H, *B;
        long long cap;
        double m, cl2, cs2;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const SpM2 Hi = sp_m2_inv(sp_m2_load(P.H, P.cap, i));
        const SpM2 A = sp_m2_mul(sp_m2_load(P.A, P.cap, i), Hi);
        sp_m2_store(P.A, P.cap, i, A);
        const SpM2 At = sp_m2_trans(A);
        const SpM2 G = sp_m2_mul(At, A);
        const double Pr = P.cl2 * (sp_m2_det(A) - 1.0);
        const SpM2 M = sp_m2_add(sp_m2_scale(Pr, sp_m2_inv(At)), sp_m2_mul(sp_m2_scale(P.cs2, A), sp_m2_dev(G, nullptr)));
        sp_m2_store(P.B, P.cap, i, sp_m2_mul(sp_m2_scale(P.m, M), Hi));
    }
};
// pull!  rod.jl:162-166
struct URodPull {
    struct Params {
        const double* X1;
        double* fy;
        double X1_min, add;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.X1[i] > P.X1_min) P.fy[i] += P.add;
    }
};
// update_v!  rod.jl:168-174
struct URodUpdateV {
    struct Params {
        WV3 v;
        RV3 f;
        const double* X1;
        double hdt, m, X1_clamp;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        double vx = P.v.x[i] + P.hdt * P.f.x[i] / P.m, vy = P.v.y[i] + P.hdt * P.f.y[i] / P.m,
               vz = P.v.z[i] + P.hdt * P.f.z[i] / P.m;
        if (P.X1[i] < P.X1_clamp) vx = vy = vz = 0.0;  // Dirichlet boundary condition
        P.v.x[i] = vx; P.v.y[i] = vy; P.v.z[i] = vz;
    }
};
// update_x!  rod.jl:176-183
struct URodUpdateX {
    struct Params {
        WV3 x;
        RV3 v;
        double *A, *H;
        WV3 f;
        double* e;
        long long cap;
        double dt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        P.x.x[i] += P.dt * P.v.x[i]; P.x.y[i] += P.dt * P.v.y[i]; P.x.z[i] += P.dt * P.v.z[i];
        for (int c = 0; c < 9; c++) {
            P.A[(size_t)c * P.cap + i] = 0.0;
            P.H[(size_t)c * P.cap + i] = 0.0;
        }
        P.f.x[i] = 0.0; P.f.y[i] = 0.0; P.f.z[i] = 0.0;
        P.e[i] = 0.0;
    }
};
// find_stress!  SHTC/ldc.jl:118-121
struct UShtcFindStress {
    struct Params {
        const double *A, *rho;
        double* stress;
        long long cap;
        double cl2, cs2, rho_ref;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const SpM3 A = sp_m3_load(P.A, P.cap, i);
        const SpM3 G = sp_m3_tmul(A, A);
        const double rho = P.rho[i];
        SpM3 S = sp_m3_mul(sp_m3_scale(P.cs2 * rho, G), sp_m3_dev(G));
        const double iso = P.cl2 * (rho - P.rho_ref);
        S.a[0] += iso; S.a[4] += iso; S.a[8] += iso;
        sp_m3_store(P.stress, P.cap, i, S);
    }
};
// relax_A!  SHTC/ldc.jl:102-116: one RK4 step of dA/dt = -3/tau*A*dev(A'*A)
struct UShtcRelaxA {
    struct Params {
        double* A;
        long long cap;
        double dt, m3_over_tau;  // -3/tau
    };
    __device__ static __forceinline__ SpM3 f(const SpM3& B, double c) {
        return sp_m3_mul(sp_m3_scale(c, B), sp_m3_dev(sp_m3_tmul(B, B)));
    }
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const SpM3 A = sp_m3_load(P.A, P.cap, i);
        const double dt = P.dt, c = P.m3_over_tau;
        SpM3 K = f(A, c);
        SpM3 R = sp_m3_add(A, sp_m3_over(sp_m3_scale(dt, K), 6.0));
        K = f(sp_m3_add(A, sp_m3_over(sp_m3_scale(dt, K), 2.0)), c);
        R = sp_m3_add(R, sp_m3_over(sp_m3_scale(dt, K), 3.0));
        K = f(sp_m3_add(A, sp_m3_over(sp_m3_scale(dt, K), 2.0)), c);
        R = sp_m3_add(R, sp_m3_over(sp_m3_scale(dt, K), 3.0));
        K = f(sp_m3_add(A, sp_m3_scale(dt, K)), c);
        R = sp_m3_add(R, sp_m3_over(sp_m3_scale(dt, K), 6.0));
        sp_m3_store(P.A, P.cap, i, R);
    }
};
// move!  SHTC/ldc.jl:129-133
struct UShtcMove {
    struct Params {
        WV3 x;
        RV3 v;
        const double* type;
        double dt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            P.x.x[i] += P.v.x[i] * P.dt; P.x.y[i] += P.v.y[i] * P.dt; P.x.z[i] += P.v.z[i] * P.dt;
        }
    }
};
// update_A!  SHTC/beryllium.jl:148-151
struct UBeUpdateA {
    struct Params {
        double *A, *T, *L;
        long long cap;
        double hdt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const SpM2 L = sp_m2_mul(sp_m2_load(P.L, P.cap, i), sp_m2_inv(sp_m2_load(P.T, P.cap, i)));
        sp_m2_store(P.L, P.cap, i, L);
        const SpM2 hL = sp_m2_scale(P.hdt, L);
        const SpM2 minus{1.0 - hL.a11, 0.0 - hL.a21, 0.0 - hL.a12, 1.0 - hL.a22};
        const SpM2 plus{1.0 + hL.a11, 0.0 + hL.a21, 0.0 + hL.a12, 1.0 + hL.a22};
        const SpM2 A = sp_m2_mul(sp_m2_mul(sp_m2_load(P.A, P.cap, i), minus), sp_m2_inv(plus));
        const double a33 = P.A[8 * P.cap + i];
        sp_m2_store(P.A, P.cap, i, A);
        P.A[8 * P.cap + i] = a33;
    }
};
// find_T!  SHTC/beryllium.jl:160-164
struct UBeFindT {
    struct Params {
        const double* A;
        double *T, *P;
        const double* J;
        long long cap;
        double rho0, c02, cs2;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const SpM2 A = sp_m2_load(P.A, P.cap, i);
        const double a33 = P.A[8 * P.cap + i], g33 = a33 * a33;
        const SpM2 G = sp_m2_mul(sp_m2_trans(A), A);
        const double J = P.J[i];
        const double Pr = 0.5 * P.rho0 * P.c02 * ((1.0 - 1.0 / J) / (J * J) + log(J) / J);
        P.P[i] = Pr;
        const double tr = 1.0 / 3.0 * (G.a11 + G.a22 + g33);
        const SpM2 D{G.a11 - tr, G.a21, G.a12, G.a22 - tr};
        const SpM2 S = sp_m2_mul(sp_m2_mul(sp_m2_scale(P.cs2, G), D), sp_m2_inv(sp_m2_load(P.T, P.cap, i)));
        const double iso = Pr / P.rho0;
        sp_m2_store(P.T, P.cap, i, SpM2{iso - S.a11, -S.a21, -S.a12, iso - S.a22});
        P.T[8 * P.cap + i] = iso - P.cs2 * g33 * (g33 - tr);
    }
};
// reset!  SHTC/beryllium.jl:177-184
struct UBeReset {
    struct Params {
        WV3 f;
        double *L, *T, *J, *Kf;
        const double *J0, *K0;
        long long cap;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        P.f.x[i] = 0.0; P.f.y[i] = 0.0; P.f.z[i] = 0.0;
        for (int c = 0; c < 9; c++) {
            P.L[(size_t)c * P.cap + i] = 0.0;
            P.T[(size_t)c * P.cap + i] = 0.0;
        }
        P.J[i] = P.J0[i];
        P.Kf[i] = P.K0[i];
    }
};
// update_v!  SHTC/beryllium.jl:132-134
struct UBeUpdateV {
    struct Params {
        WV3 v;
        RV3 f;
        const double* m;
        double hdt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const double m = P.m[i];
        P.v.x[i] += P.hdt * P.f.x[i] / m; P.v.y[i] += P.hdt * P.f.y[i] / m; P.v.z[i] += P.hdt * P.f.z[i] / m;
    }
};
// update_A!  SHTC/twist3d.jl:143-146
struct UTwUpdateA {
    struct Params {
        double *A, *T, *L;
        long long cap;
        double hdt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const SpM3 L = sp_m3_mul(sp_m3_load(P.L, P.cap, i), sp_m3_inv(sp_m3_load(P.T, P.cap, i)));
        sp_m3_store(P.L, P.cap, i, L);
        SpM3 minus, plus;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const double e = (k % 4 == 0) ? 1.0 : 0.0, hl = P.hdt * L.a[k];
            minus.a[k] = e - hl;
            plus.a[k] = e + hl;
        }
        sp_m3_store(P.A, P.cap, i, sp_m3_mul(sp_m3_mul(sp_m3_load(P.A, P.cap, i), minus), sp_m3_inv(plus)));
    }
};
// find_T!  SHTC/twist3d.jl:155-161
struct UTwFindT {
    struct Params {
        const double* A;
        double *T, *P;
        const double* J;
        long long cap;
        double rho0, c02, cs2;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const SpM3 F = sp_m3_inv(sp_m3_load(P.A, P.cap, i));
        SpM3 BmI;  // F*F' - I
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int r = 0; r < 3; r++)
                BmI.a[r + 3 * j] = (F.a[r] * F.a[j] + F.a[r + 3] * F.a[j + 3] + F.a[r + 6] * F.a[j + 6]) - (r == j ? 1.0 : 0.0);
        const double detF = 1.0 / P.J[i];
        const double Pr = -P.rho0 * P.c02 * (detF * detF) * (detF - 1.0);
        P.P[i] = Pr;
        const SpM3 S = sp_m3_mul(sp_m3_scale(P.cs2, BmI), sp_m3_inv(sp_m3_load(P.T, P.cap, i)));
        const double iso = -Pr / P.rho0;
        SpM3 T;
#pragma unroll
        for (int k = 0; k < 9; k++) T.a[k] = ((k % 4 == 0) ? iso : 0.0) - S.a[k];
        sp_m3_store(P.T, P.cap, i, T);
    }
};
// update_v!  SHTC/twist3d.jl:125-129
struct UTwUpdateV {
    struct Params {
        const double* z;
        WV3 v;
        RV3 f;
        const double* m;
        double hdt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.z[i] > 0.0) {
            const double m = P.m[i];
            P.v.x[i] += P.hdt * P.f.x[i] / m; P.v.y[i] += P.hdt * P.f.y[i] / m; P.v.z[i] += P.hdt * P.f.z[i] / m;
        }
    }
};
// find_T!  SHTC/taco.jl:148-152 (subinv: zeros outside the in-plane block)
struct UTaFindT {
    struct Params {
        const double* A;
        double *T, *P;
        const double* rho;
        long long cap;
        double rho0, c02, cs2;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const SpM3 A = sp_m3_load(P.A, P.cap, i);
        const SpM3 G = sp_m3_tmul(A, A);
        const SpM3 GD = sp_m3_mul(sp_m3_scale(P.cs2, G), sp_m3_dev(G));
        const double rho = P.rho[i];
        const double Pr = P.c02 * (rho - P.rho0) * P.rho0 / rho;
        P.P[i] = Pr;
        const SpM2 si = sp_m2_inv(sp_m2_load(P.T, P.cap, i));
        const double iso = -Pr / (rho * rho);
        SpM3 S;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            S.a[r] = GD.a[r] * si.a11 + GD.a[r + 3] * si.a21;
            S.a[r + 3] = GD.a[r] * si.a12 + GD.a[r + 3] * si.a22;
            S.a[r + 6] = 0.0;
        }
        S.a[0] += iso; S.a[4] += iso; S.a[8] += iso;
        sp_m3_store(P.T, P.cap, i, S);
    }
};
// update_v!  SHTC/taco.jl:108-114 with vexact :39-42
struct UTaUpdateV {
    struct Params {
        RV3 x;
        WV3 v;
        RV3 f;
        const double *m, *type;
        double hdt, R1, R2, omega;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            const double m = P.m[i];
            P.v.x[i] += P.hdt * P.f.x[i] / m; P.v.y[i] += P.hdt * P.f.y[i] / m; P.v.z[i] += P.hdt * P.f.z[i] / m;
        } else {
            const double x = P.x.x[i], y = P.x.y[i], z = P.x.z[i];
            const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
            const double sc = P.R2 / r * (r / P.R1 - P.R1 / r) / (P.R2 / P.R1 - P.R1 / P.R2);
            P.v.x[i] = sc * (-P.omega * y); P.v.y[i] = sc * (P.omega * x); P.v.z[i] = sc * 0.0;
        }
    }
};
// update_x!  SHTC/taco.jl:116-126
struct UTaUpdateX {
    struct Params {
        WV3 x;
        RV3 v, x0;
        const double* type;
        double hdt, cw, sw, outer;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        const double t = P.type[i];
        if (t == 0.0) {
            P.x.x[i] += P.hdt * P.v.x[i]; P.x.y[i] += P.hdt * P.v.y[i]; P.x.z[i] += P.hdt * P.v.z[i];
        } else if (t == P.outer) {
            const double a = P.x0.x[i], b = P.x0.y[i];
            P.x.x[i] = __dsub_rn(__dmul_rn(a, P.cw), __dmul_rn(b, P.sw));
            P.x.y[i] = __dadd_rn(__dmul_rn(a, P.sw), __dmul_rn(b, P.cw));
            P.x.z[i] = 0.0;
        }
    }
};
// find_pressure!  test_collision_2d.jl:71-73
struct UPressureFromRho {
    struct Params {
        const double *rho, *rho0;
        double* P;
        double c2;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) { P.P[i] = P.c2 * (P.rho[i] - P.rho0[i]); }
};
// reset_a! / reset_rho!  test_collision_2d.jl:80-86
struct UFill {
    struct Params {
        double* f;
        long long cap;
        int ncomp;
        double value;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        for (int c = 0; c < P.ncomp; c++) P.f[(size_t)c * P.cap + i] = P.value;
    }
};
// move!  test_collision_2d.jl:88-90
struct UAdvect {
    struct Params {
        WV3 x;
        RV3 v;
        double dt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        P.x.x[i] += P.dt * P.v.x[i]; P.x.y[i] += P.dt * P.v.y[i]; P.x.z[i] += P.dt * P.v.z[i];
    }
};
// accelerate!  test_collision_2d.jl:92-94
struct UKick {
    struct Params {
        WV3 v;
        RV3 a;
        double hdt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        P.v.x[i] += P.hdt * P.a.x[i]; P.v.y[i] += P.hdt * P.a.y[i]; P.v.z[i] += P.hdt * P.a.z[i];
    }
};
// initialize!  collapse_dry_implicit.jl:118-126
struct UIsphInitialize {
    struct Params {
        WV3 x, v;
        double *div, *L, *lambda;
        const double* type;
        double dt, gx, gy, gz;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            P.x.x[i] += P.dt * P.v.x[i]; P.x.y[i] += P.dt * P.v.y[i]; P.x.z[i] += P.dt * P.v.z[i];
            P.v.x[i] += P.dt * P.gx; P.v.y[i] += P.dt * P.gy; P.v.z[i] += P.dt * P.gz;
        }
        P.div[i] = 0.0;
        P.L[i] = 0.0;
        P.lambda[i] = 1.0;
    }
};
// projection_vector  collapse_dry_implicit.jl:165-167
struct UIsphProjectionVector {
    struct Params {
        const double* div;
        double* b;
        double neg_h2, dt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) { P.b[i] = P.neg_h2 * P.div[i] / P.dt; }
};
// accelerate!  collapse_dry_implicit.jl:136-141
struct UIsphAccelerate {
    struct Params {
        WV3 v, Dv;
        const double* type;
        double dt;
    };
    __device__ static __forceinline__ void apply(const Params& P, int i) {
        if (P.type[i] == 0.0) {
            P.v.x[i] += P.dt * P.Dv.x[i]; P.v.y[i] += P.dt * P.Dv.y[i]; P.v.z[i] += P.dt * P.Dv.z[i];
        }
        P.Dv.x[i] = 0.0; P.Dv.y[i] = 0.0; P.Dv.z[i] = 0.0;
    }
};
