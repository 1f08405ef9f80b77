// sp_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.
//
// A C++17/OpenMP restatement of the reference's algorithm for the hot path of
// SmoothedParticles.jl v0.2.0 (pure Julia; the reference cannot run in this image because
// no Julia runtime exists here or on the GPU box).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / `--impl reference` leg may load this library; the product
// (smoothedparticles.jl_b200/) never does.
//
// PARITY UNPINNED against output of the Julia package itself (no Julia runtime here or on the GPU box, no network, and
// the reference ships no golden vectors for this path) — what follows is the strongest pin obtainable without it.
// PARITY PINNING.  The reference ships no golden vectors for this path.  What pins this
// restatement to the Julia code is (1) the reference's own assertions, ported 1:1 and run
// against this file in tests/test_oracle_pins.py: tests/test_kernels.jl:20-61 (kernel
// known-answer properties) and tests/test_collision_2d.jl:118-149 (particle count constant,
// energy growth < 1e-2 over 4 167 Verlet steps); (2) an independent brute-force O(N^2) /
// scipy cKDTree neighbour check; (3) O(N^2) numpy evaluations of the example closures' formulas
// for the operators no reference test exercises (symplectic / cylinder / rod scripts,
// assemble_matrix), exact integer arithmetic for rev_add, scipy's CG for the CG restatement
// (tests/test_symplectic_cpu.py, test_cylinder_cpu.py, test_rod_cpu.py, test_oracle_pins.py).
// No output of the Julia reference itself could be generated, and that limitation is stated
// in DESIGN.md.
//
// Every function cites the reference file:line it follows (paths relative to the
// reference root).  Compile with -ffp-contract=off and without fast-math so each
// arithmetic operation is individually rounded, as Julia does outside @fastmath.
//
// Particle storage mirrors the reference's array of mutable structs: one record of NSLOT
// Float64 per particle; an operator's "field binding" is the offset of a struct field in
// the record (the C analogue of p.rho, p.Dv ...).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/sp_b200.h"  // operator / kernel ids only

// Float64 slots per particle record.  The default (20 = 160 B) is what bench.py's cpu_baseline runs on and fits every
// scalar/vector example; oracle/Makefile also builds a wide variant (-DNSLOT=48) for the tensor-field example rod.jl.
#ifndef NSLOT
#define NSLOT 20
#endif

namespace {

struct Particle {
    double f[NSLOT];
};

struct OSys {
    double h;
    double lo[3], hi[3];  // domain == its bounding Box (structs.jl:63,87)
    int64_t key_phase[3], key_lim[3], key_max;
    std::vector<int64_t> key_diff;
    std::vector<Particle> particles;
    // cell_list as CSR: cell k (1-based) = entries[cell_start[k-1] .. cell_start[k]), 1-based particle
    // indices in DESCENDING order == the non-zero prefix of Cell.entries after add_index! (core.jl:26-41).
    std::vector<int64_t> cell_start;
    std::vector<int64_t> entries;
    bool have_cells = false;
    int64_t n_removed = 0;
    std::string err;
};

// ------------------------------------------------------------------ kernels.jl
// @fastmath in the reference: only a tolerance (not bits) is meaningful against these.
inline double pw2(double x) { return x * x; }
inline double pw3(double x) { return x * x * x; }
inline double pw4(double x) { return (x * x) * (x * x); }
inline double pos(double x) { return x > 0.0 ? x : 0.0; }  // kernels.jl:3-5

double spline23(double h, double r) {  // kernels.jl:14-24
    double x = r / h;
    if (x < 0.5) return 1.8189136353359467 * (1.0 - 6.0 * pw2(x) + 6.0 * pw3(x)) / pw2(h);
    else if (x < 1.0) return 3.6378272706718935 * pw3(1.0 - x) / pw2(h);
    return 0.0;
}
double Dspline23(double h, double r) {  // kernels.jl:33-42
    double x = r / h;
    if (x < 0.5) return -10.91348181201568 * (2.0 * x - 3.0 * pw2(x)) / pw3(h);
    else if (x < 1.0) return -10.91348181201568 * pw2(1.0 - x) / pw3(h);
    return 0.0;
}
double rDspline23(double h, double r) {  // kernels.jl:51-60
    double x = r / h;
    if (x < 0.5) return -10.91348181201568 * (2.0 - 3.0 * x) / pw4(h);
    else if (x < 1.0) return -10.91348181201568 * pw2(1.0 - x) / (x * pw4(h));
    return 0.0;
}
double spline24(double h, double r) {  // kernels.jl:69-72
    double x = r / h;
    return 6.222175110452539 * (pw4(pos(1.0 - x)) - 5 * pw4(pos(0.6 - x)) + 10 * pw4(pos(0.2 - x))) / pw2(h);
}
double Dspline24(double h, double r) {  // kernels.jl:81-84
    double x = r / h;
    return -24.888700441810155 * (pw3(pos(1.0 - x)) - 5 * pw3(pos(0.6 - x)) + 10 * pw3(pos(0.2 - x))) / pw3(h);
}
double rDspline24(double h, double r) {  // kernels.jl:93-99
    double x = r / h;
    if (x > 0.2) return -24.888700441810155 * (pw3(pos(1.0 - x)) - 5 * pw3(pos(0.6 - x))) / (x * pw4(h));
    return -24.888700441810155 * (1.2 - 6.0 * pw2(x)) / pw4(h);
}
double wendland2(double h, double r) {  // kernels.jl:108-115
    double x = r / h;
    if (x > 1.0) return 0.0;
    return 2.228169203286535 * pw4(1.0 - x) * (1.0 + 4.0 * x) / pw2(h);
}
double Dwendland2(double h, double r) {  // kernels.jl:124-131
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -44.563384065730695 * x * pw3(1.0 - x) / pw3(h);
}
double rDwendland2(double h, double r) {  // kernels.jl:140-147
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -44.563384065730695 * pw3(1.0 - x) / pw4(h);
}
double wendland3(double h, double r) {  // kernels.jl:156-163
    double x = r / h;
    if (x > 1.0) return 0.0;
    return 3.3422538049298023 * pw4(1.0 - x) * (1.0 + 4.0 * x) / pw3(h);
}
double Dwendland3(double h, double r) {  // kernels.jl:172-179
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -66.84507609859604 * x * pw3(1.0 - x) / pw4(h);
}
double rDwendland3(double h, double r) {  // kernels.jl:188-195
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -66.84507609859604 * pw3(1.0 - x) / (pw4(h) * h);
}
double DDwendland3(double h, double r) {  // kernels.jl:197-204
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -66.84507609859604 * ((1.0 - 4.0 * x) * pw2(1.0 - x)) / (pw4(h) * h);
}
double wendland1(double h, double r) {  // kernels.jl:206-212
    double x = r / h;
    if (x > 1.0) return 0.0;
    return 1.5 * pw4(1.0 - x) * (1.0 + 4.0 * x) / h;
}
double Dwendland1(double h, double r) {  // kernels.jl:214-220
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -30.0 * x * pw3(1.0 - x) / pw2(h);
}
double rDwendland1(double h, double r) {  // kernels.jl:222-228
    double x = r / h;
    if (x > 1.0) return 0.0;
    return -30.0 * pw3(1.0 - x) / pw3(h);
}

double kernel_eval(int kernel, int kfun, double h, double r) {
    switch (kernel) {
        case SP_KERNEL_WENDLAND1:
            return kfun == SP_KFUN_W ? wendland1(h, r) : kfun == SP_KFUN_DW ? Dwendland1(h, r) : rDwendland1(h, r);
        case SP_KERNEL_WENDLAND2:
            return kfun == SP_KFUN_W ? wendland2(h, r) : kfun == SP_KFUN_DW ? Dwendland2(h, r) : rDwendland2(h, r);
        case SP_KERNEL_WENDLAND3:
            return kfun == SP_KFUN_W    ? wendland3(h, r)
                   : kfun == SP_KFUN_DW ? Dwendland3(h, r)
                   : kfun == SP_KFUN_DDW ? DDwendland3(h, r)
                                         : rDwendland3(h, r);
        case SP_KERNEL_SPLINE23:
            return kfun == SP_KFUN_W ? spline23(h, r) : kfun == SP_KFUN_DW ? Dspline23(h, r) : rDspline23(h, r);
        case SP_KERNEL_SPLINE24:
            return kfun == SP_KFUN_W ? spline24(h, r) : kfun == SP_KFUN_DW ? Dspline24(h, r) : rDspline24(h, r);
    }
    return std::numeric_limits<double>::quiet_NaN();
}

// ------------------------------------------------------------------ structs.jl
// ParticleSystem constructor, structs.jl:57-91
void init_keys(OSys& s) {
    for (int a = 0; a < 3; a++) {
        s.key_phase[a] = (int64_t)std::floor(s.lo[a] / s.h);                        // :66
        s.key_lim[a] = (int64_t)std::floor(s.hi[a] / s.h) - s.key_phase[a] + 1;    // :67
    }
    s.key_max = s.key_lim[0] * s.key_lim[1] * s.key_lim[2];  // :68
    s.key_diff.clear();
    if (s.key_lim[2] == 1) {  // :70  2-D
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++) s.key_diff.push_back(di + s.key_lim[0] * dj);  // :73-75
    } else {
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
                for (int dk = -1; dk <= 1; dk++)
                    s.key_diff.push_back(di + s.key_lim[0] * (dj + s.key_lim[1] * dk));  // :79-81
    }
}

// find_key, structs.jl:97-106.  Int64(floor(x/h)) throws on NaN/Inf -> -1.
inline int64_t find_key(const OSys& s, const double* x) {
    double q0 = std::floor(x[0] / s.h), q1 = std::floor(x[1] / s.h), q2 = std::floor(x[2] / s.h);
    if (!std::isfinite(q0) || !std::isfinite(q1) || !std::isfinite(q2)) return -1;
    if (std::fabs(q0) > 9.0e18 || std::fabs(q1) > 9.0e18 || std::fabs(q2) > 9.0e18) return -1;
    int64_t i = 1 + (int64_t)q0 - s.key_phase[0];
    int64_t j = 1 + (int64_t)q1 - s.key_phase[1];
    int64_t k = 1 + (int64_t)q2 - s.key_phase[2];
    return i + s.key_lim[0] * (j - 1) + s.key_lim[0] * s.key_lim[1] * (k - 1);
}

// is_inside(x, Box), geometry.jl:24-30 — closed intervals, NaN compares false => outside.
inline bool is_inside_box(const OSys& s, const double* x) {
    return s.lo[0] <= x[0] && x[0] <= s.hi[0] && s.lo[1] <= x[1] && x[1] <= s.hi[1] && s.lo[2] <= x[2] &&
           x[2] <= s.hi[2];
}

// ------------------------------------------------------------------ core.jl
// dist, core.jl:8-10 -> norm/dot algebra.jl:49-60: sqrt((dx*dx + dy*dy) + dz*dz), un-fused.
inline double dist3(const double* a, const double* b, double* d) {
    d[0] = a[0] - b[0];
    d[1] = a[1] - b[1];
    d[2] = a[2] - b[2];
    return std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
}

// Removal of out-of-domain particles, core.jl:63-81 (literal, serial).
// removal_cell.entries ends up sorted DESCENDING (add_index!, core.jl:26-41); the i-th victim slot
// receives the CURRENT particles[end+1-i]; then the vector is truncated.
int64_t remove_outside(OSys& s) {
    const int64_t N = (int64_t)s.particles.size();
    std::vector<int64_t> victims;  // 1-based
    for (int64_t i = N; i >= 1; i--)
        if (!is_inside_box(s, s.particles[i - 1].f)) victims.push_back(i);  // descending
    int64_t i = 1;
    while (i <= (int64_t)victims.size()) {
        s.particles[victims[i - 1] - 1] = s.particles[N + 1 - i - 1];  // :74
        i++;
    }
    if (i > 1) s.particles.resize(N + 1 - i);  // :77-79
    return i - 1;
}

// create_cell_list!, core.jl:51-90.  The observable result (each cell's member indices in descending
// order) is independent of thread interleaving; it is produced here by a stable counting sort.
void create_cell_list(OSys& s) {
    s.n_removed += remove_outside(s);
    const int64_t N = (int64_t)s.particles.size();
    std::vector<int64_t> keys(N);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) keys[i] = find_key(s, s.particles[i].f);  // :86
    s.cell_start.assign(s.key_max + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) {
#pragma omp atomic
        s.cell_start[keys[i]]++;  // keys are 1-based: count of cell k at [k]
    }
    // exclusive scan: cell_start[k-1] = first slot of cell k
    int64_t run = 0;
    for (int64_t k = 1; k <= s.key_max; k++) {
        int64_t c = s.cell_start[k];
        s.cell_start[k - 1] = run;
        run += c;
    }
    s.cell_start[s.key_max] = run;
    s.entries.resize(N);
    std::vector<int64_t> fill(s.cell_start.begin(), s.cell_start.end() - 1);
#pragma omp parallel for schedule(static)
    for (int64_t i = 1; i <= N; i++) {
        int64_t slot;
#pragma omp atomic capture
        slot = fill[keys[i - 1] - 1]++;
        s.entries[slot] = i;
    }
    // add_index! keeps every cell sorted DESCENDING whatever the insertion order (core.jl:32-37)
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t k = 1; k <= s.key_max; k++)
        if (s.cell_start[k] - s.cell_start[k - 1] > 1)
            std::sort(s.entries.begin() + s.cell_start[k - 1], s.entries.begin() + s.cell_start[k],
                      std::greater<int64_t>());
    s.have_cells = true;
}

// Literal restatement of the insertion path (find_vacation!/add_index!, core.jl:13-41) used only to
// cross-check create_cell_list() in the tests: cells are growable zero-padded vectors.
void create_cell_list_literal(OSys& s, std::vector<std::vector<int64_t>>& cells) {
    cells.assign(s.key_max, {});
    const int64_t N = (int64_t)s.particles.size();
    for (int64_t i = 1; i <= N; i++) {
        int64_t key = find_key(s, s.particles[i - 1].f);
        std::vector<int64_t>& e = cells[key - 1];
        size_t ind = 0;
        while (ind < e.size() && e[ind] != 0) ind++;  // find_vacation!
        if (ind == e.size()) e.resize(ind + 1);
        e[ind] = i;
        while (ind > 0 && e[ind - 1] < e[ind]) {  // reorder
            std::swap(e[ind], e[ind - 1]);
            ind--;
        }
    }
}

// _apply_binary!, core.jl:94-112, for the particle with 1-based index ip.
template <class Action>
inline void apply_binary_one(OSys& s, int64_t ip, Action&& action) {
    Particle& p = s.particles[ip - 1];
    int64_t key = find_key(s, p.f);  // :95
    for (int64_t dkey : s.key_diff) {  // :96
        int64_t nk = key + dkey;
        if (1 <= nk && nk <= s.key_max) {  // :98
            for (int64_t e = s.cell_start[nk - 1]; e < s.cell_start[nk]; e++) {
                int64_t j = s.entries[e];
                const Particle& q = s.particles[j - 1];
                double d[3];
                double r = dist3(p.f, q.f, d);      // :104
                if (r > s.h || j == ip) continue;  // :105  (p == q is object identity)
                action(p, q, d, r);
            }
        }
    }
}

// apply_binary!, core.jl:125-129
template <class Action>
void apply_binary(OSys& s, Action&& action) {
    const int64_t N = (int64_t)s.particles.size();
#pragma omp parallel for schedule(static)
    for (int64_t i = 1; i <= N; i++) apply_binary_one(s, i, action);
}
// apply_unary!, core.jl:138-142
template <class Action>
void apply_unary(OSys& s, Action&& action) {
    const int64_t N = (int64_t)s.particles.size();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) action(s.particles[i]);
}

typedef double (*kfn)(double, double);
kfn pick_rD(int kernel) {
    switch (kernel) {
        case SP_KERNEL_WENDLAND1: return rDwendland1;
        case SP_KERNEL_WENDLAND2: return rDwendland2;
        case SP_KERNEL_WENDLAND3: return rDwendland3;
        case SP_KERNEL_SPLINE23: return rDspline23;
        case SP_KERNEL_SPLINE24: return rDspline24;
    }
    return nullptr;
}
kfn pick_w(int kernel) {
    switch (kernel) {
        case SP_KERNEL_WENDLAND1: return wendland1;
        case SP_KERNEL_WENDLAND2: return wendland2;
        case SP_KERNEL_WENDLAND3: return wendland3;
        case SP_KERNEL_SPLINE23: return spline23;
        case SP_KERNEL_SPLINE24: return spline24;
    }
    return nullptr;
}

inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }  // algebra.jl:49-51

// ---- examples/utils/FixPA.jl:11-42: reversible fixed-point addition.  FixPA_eps = 1/2^30, so x/FixPA_eps is the
// exact product x*2^30; Julia's round() is round-half-to-even (nearbyint in the default rounding mode).
inline int64_t fixpa_nom(double x) { return (int64_t)std::nearbyint(x / (1.0 / 1073741824.0)); }  // :19-21
inline double rev_add(double x, double y) { return (1.0 / 1073741824.0) * (double)(fixpa_nom(x) + fixpa_nom(y)); }  // :28-30
// s^4 with an integer literal exponent is Base.pow_body(x, 4) (Julia >= 1.8: power by squaring with the low parts
// of the two squarings carried along); written out for n = 4.  Older Julia: (s*s)*(s*s), at most 1 ulp away.
inline double julia_pow4(double x) {
    double x2 = x * x, lo2 = std::fma(x, x, -x2);
    double err = x2 * 2 * lo2;
    double x4 = x2 * x2, lo4 = std::fma(x2, x2, -x4);
    lo4 += err;
    return (std::isfinite(x4) && std::isfinite(lo4)) ? x4 + lo4 : x4;
}

// ---- examples/rod.jl:44-85: the script's own 2-D matrix helpers.  A RealMatrix field occupies 9 slots in Julia's
// column-major order (slot c = (i-1) + 3*(j-1)); rod.jl's outer/inv/trans/dev only ever populate M[1:2,1:2]
// (dev also sets M[3,3], which every product here multiplies by a zero), so the in-plane block is what is computed.
struct M2 {
    double a11, a21, a12, a22;
};
inline M2 m2_load(const double* f) { return M2{f[0], f[1], f[3], f[4]}; }
inline void m2_store(double* f, const M2& a) {
    f[0] = a.a11; f[1] = a.a21; f[3] = a.a12; f[4] = a.a22;
    f[2] = f[5] = f[6] = f[7] = f[8] = 0.0;
}
inline double m2_det(const M2& a) { return a.a11 * a.a22 - a.a12 * a.a21; }  // :52-54
inline M2 m2_inv(const M2& a) {                                               // :56-63
    double idet = 1.0 / m2_det(a);
    return M2{+idet * a.a22, -idet * a.a21, -idet * a.a12, +idet * a.a11};
}
inline M2 m2_trans(const M2& a) { return M2{a.a11, a.a12, a.a21, a.a22}; }  // :65-71
inline M2 m2_mul(const M2& a, const M2& b) {  // SMatrix product restricted to the block (third terms are 0*0)
    return M2{a.a11 * b.a11 + a.a12 * b.a21, a.a21 * b.a11 + a.a22 * b.a21, a.a11 * b.a12 + a.a12 * b.a22,
              a.a21 * b.a12 + a.a22 * b.a22};
}
inline void m2_vec(const M2& a, const double* x, double* y) {
    y[0] = a.a11 * x[0] + a.a12 * x[1];
    y[1] = a.a21 * x[0] + a.a22 * x[1];
}
inline M2 m2_scale(double c, const M2& a) { return M2{c * a.a11, c * a.a21, c * a.a12, c * a.a22}; }
inline M2 m2_add(const M2& a, const M2& b) { return M2{a.a11 + b.a11, a.a21 + b.a21, a.a12 + b.a12, a.a22 + b.a22}; }
// dev :73-80; lam_out receives lambda so that the caller can form the [3,3] entry 1 - lambda
inline M2 m2_dev(const M2& g, double* lam_out) {
    double lam = 1.0 / 3.0 * (g.a11 + g.a22 + 1.0);
    if (lam_out) *lam_out = lam;
    return M2{g.a11 - lam, g.a21, g.a12, g.a22 - lam};
}

// ---- full 3x3 RealMatrix algebra for examples/SHTC/*.jl: 9 slots in Julia's column-major order, a[i + 3*j] = M[i+1,j+1].
// StaticArrays' SMatrix product is restated as the plain sum over the inner index (the package is neither vendored nor
// pinned; whether it contracts to FMA is not observable from the reference — tolerance-level either way).
struct M3 {
    double a[9];
};
inline M3 m3_load(const double* f) {
    M3 m;
    for (int c = 0; c < 9; c++) m.a[c] = f[c];
    return m;
}
inline void m3_store(double* f, const M3& m) {
    for (int c = 0; c < 9; c++) f[c] = m.a[c];
}
inline M3 m3_mul(const M3& A, const M3& B) {
    M3 C;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) C.a[i + 3 * j] = A.a[i] * B.a[3 * j] + A.a[i + 3] * B.a[1 + 3 * j] + A.a[i + 6] * B.a[2 + 3 * j];
    return C;
}
inline M3 m3_tmul(const M3& A, const M3& B) {  // A'*B
    M3 C;
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++)
            C.a[i + 3 * j] = A.a[3 * i] * B.a[3 * j] + A.a[1 + 3 * i] * B.a[1 + 3 * j] + A.a[2 + 3 * i] * B.a[2 + 3 * j];
    return C;
}
inline M3 m3_scale(double c, const M3& A) {
    M3 C;
    for (int k = 0; k < 9; k++) C.a[k] = c * A.a[k];
    return C;
}
inline M3 m3_add(const M3& A, const M3& B) {
    M3 C;
    for (int k = 0; k < 9; k++) C.a[k] = A.a[k] + B.a[k];
    return C;
}
inline M3 m3_dev(const M3& G) {  // deviatoric, ldc.jl:84-86: G - 1/3*(G11 + G22 + G33)*MAT1
    const double lam = 1.0 / 3.0 * (G.a[0] + G.a[4] + G.a[8]);
    M3 C = G;
    C.a[0] = G.a[0] - lam;
    C.a[4] = G.a[4] - lam;
    C.a[8] = G.a[8] - lam;
    return C;
}
inline M3 shtc_relax_f(const M3& A, double tau) {  // ldc.jl:102-104
    return m3_mul(m3_scale(-3.0 / tau, A), m3_dev(m3_tmul(A, A)));
}

// ---- examples/SHTC/beryllium.jl:44-52: the "structural" kernels (strict x < 1, plain arithmetic restatement of the
// script's @fastmath expressions) and :91-98 the script's 2-D inverse, which sets [3,3] = 1
inline double wendland2h(double h, double r) {
    const double x = r / h;
    return x < 1.0 ? 14.0 * pw3(1.0 - x) * (14.0 * pw2(x) - 3.0 * x - 1.0) / (M_PI * pw2(h)) : 0.0;
}
inline double rDwendland2h(double h, double r) {
    const double x = r / h;
    return x < 1.0 ? 140.0 * pw2(1.0 - x) * (4.0 - 7.0 * x) / (M_PI * pw4(h)) : 0.0;
}

// ---- examples/SHTC/twist3d.jl:43-51 structural kernels in 3-D, and the general 3x3 inverse (StaticArrays.inv is not
// vendored; restated as adjugate/det like src/algebra.jl:121-158 — tolerance-level against any other formula)
inline double wendland3h(double h, double r) {
    const double x = r / h;
    return x < 1.0 ? 21.0 * pw3(1.0 - x) * (14.0 * pw2(x) - 3.0 * x - 1.0) / (M_PI * pw3(h)) : 0.0;
}
inline double rDwendland3h(double h, double r) {
    const double x = r / h;
    return x < 1.0 ? 210.0 * pw2(1.0 - x) * (4.0 - 7.0 * x) / (M_PI * (pw4(h) * h)) : 0.0;
}
inline double m3_det(const M3& A) {
    const double* a = A.a;
    return a[0] * a[4] * a[8] + a[1] * a[5] * a[6] + a[2] * a[3] * a[7] - a[6] * a[4] * a[2] - a[7] * a[5] * a[0] - a[8] * a[3] * a[1];
}
inline M3 m3_inv(const M3& A) {
    const double* a = A.a;
    const double id = 1.0 / m3_det(A);
    M3 C;  // inverse[i,j] = cofactor[j,i]/det, column-major
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
inline M3 m3_identity() {
    M3 I;
    for (int k = 0; k < 9; k++) I.a[k] = (k % 4 == 0) ? 1.0 : 0.0;
    return I;
}

// apply!, core.jl:151-161, specialised to the registered operators (the example closures).
int apply_op(OSys& s, int op, const int32_t* F, int nf, const double* P, int np, int flags) {
    auto need = [&](int f, int p) { return nf == f && np == p; };
    const bool self = (flags & SP_FLAG_SELF) != 0;
    switch (op) {
        case SP_OP_BALANCE_OF_MASS: {  // collapse_dry.jl:112-115, collapse3d.jl:87-90, cavity_flow.jl:92-94
            if (!need(4, 4)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], orho = F[2], oD = F[3];
            kfn rDw = pick_rD((int)P[0]);
            const double m = P[1], h = P[2], two_nu = P[3];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double ker = m * rDw(h, r);
                double vpq[3] = {p.f[ov] - q.f[ov], p.f[ov + 1] - q.f[ov + 1], p.f[ov + 2] - q.f[ov + 2]};
                (void)ox;
                p.f[oD] += ker * (dot3(xpq, vpq) + two_nu * (p.f[orho] - q.f[orho]));
            });
            return SP_OK;
        }
        case SP_OP_FIND_PRESSURE: {  // collapse_dry.jl:123-127, cavity_flow.jl:96-100
            if (!need(3, 4)) return SP_ERR_INVALID;
            const int orho = F[0], oD = F[1], oP = F[2];
            const double dt = P[0], c2 = P[1], rho0 = P[2], P0 = P[3];
            apply_unary(s, [=](Particle& p) {
                p.f[orho] += p.f[oD] * dt;
                p.f[oD] = 0.0;
                double pr = c2 * (p.f[orho] - rho0);
                p.f[oP] = (P0 != 0.0) ? P0 + pr : pr;
            });
            return SP_OK;
        }
        case SP_OP_INTERNAL_FORCE: {  // collapse_dry.jl:135-141
            if (!need(6, 5)) return SP_ERR_INVALID;
            const int ov = F[1], oP = F[2], orho = F[3], oDv = F[4], ot = F[5];
            kfn rDw = pick_rD((int)P[0]);
            const double m = P[1], h = P[2], mu = P[3], rho0 = P[4];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                if (p.f[ot] == 0.0) {
                    double ker = m * rDw(h, r);
                    double a = -ker * (p.f[oP] / (p.f[orho] * p.f[orho]) + q.f[oP] / (q.f[orho] * q.f[orho]));
                    for (int c = 0; c < 3; c++) p.f[oDv + c] += a * xpq[c];
                    double b = 2 * ker * mu / (rho0 * rho0);
                    for (int c = 0; c < 3; c++) p.f[oDv + c] += b * (p.f[ov + c] - q.f[ov + c]);
                }
            });
            return SP_OK;
        }
        case SP_OP_INTERNAL_FORCE_CAVITY: {  // cavity_flow.jl:102-114
            if (!need(6, 6)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], oP = F[2], orho = F[3], oDv = F[4], ot = F[5];
            const double m = P[0], h = P[1], Re = P[2], vlid = P[3], ylid = P[4], lid = P[5];
            const double eps = 0.01 * (h * h), tenth_h = 0.1 * h;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double rDk = rDwendland2(h, r);
                double vpq[3] = {p.f[ov] - q.f[ov], p.f[ov + 1] - q.f[ov + 1], p.f[ov + 2] - q.f[ov + 2]};
                if (q.f[ot] == lid) {
                    double sc = std::fabs(xpq[1]) / (tenth_h + std::fabs(p.f[ox + 1] - ylid));
                    vpq[0] = sc * (p.f[ov] - vlid * 1.0);
                    vpq[1] = sc * (p.f[ov + 1] - vlid * 0.0);
                    vpq[2] = sc * (p.f[ov + 2] - vlid * 0.0);
                }
                double a = -m * rDk * (p.f[oP] / (p.f[orho] * p.f[orho]) + q.f[oP] / (q.f[orho] * q.f[orho]));
                for (int c = 0; c < 3; c++) p.f[oDv + c] += a * xpq[c];
                double b = 8 / (Re * p.f[orho] * q.f[orho]) * m * rDk * dot3(vpq, xpq) / (r * r + eps);
                for (int c = 0; c < 3; c++) p.f[oDv + c] += b * xpq[c];
            });
            return SP_OK;
        }
        case SP_OP_MOVE: {  // collapse_dry.jl:148-153
            if (!need(4, 1)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], oDv = F[2], ot = F[3];
            const double dtm = P[0];
            apply_unary(s, [=](Particle& p) {
                p.f[oDv] = p.f[oDv + 1] = p.f[oDv + 2] = 0.0;
                if (p.f[ot] == 0.0)
                    for (int c = 0; c < 3; c++) p.f[ox + c] += dtm * p.f[ov + c];
            });
            return SP_OK;
        }
        case SP_OP_ACCELERATE: {  // collapse_dry.jl:155-159
            if (!need(3, 4)) return SP_ERR_INVALID;
            const int ov = F[0], oDv = F[1], ot = F[2];
            const double hdt = P[0], g[3] = {P[1], P[2], P[3]};
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0)
                    for (int c = 0; c < 3; c++) p.f[ov + c] += hdt * (p.f[oDv + c] + g[c]);
            });
            return SP_OK;
        }
        case SP_OP_SC_BALANCE_OF_MASS: {  // static_container.jl:102-104
            if (!need(3, 4)) return SP_ERR_INVALID;
            const int ov = F[1], orho = F[2];
            kfn rDw = pick_rD((int)P[0]);
            const double m = P[1], h = P[2], dt = P[3];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double vpq[3] = {p.f[ov] - q.f[ov], p.f[ov + 1] - q.f[ov + 1], p.f[ov + 2] - q.f[ov + 2]};
                p.f[orho] += dt * dot3(xpq, vpq) * m * rDw(h, r);
            });
            return SP_OK;
        }
        case SP_OP_SC_INTERNAL_FORCE: {  // static_container.jl:106-114, pressure(p) :68-70
            if (!need(5, 6)) return SP_ERR_INVALID;
            const int ov = F[1], orho = F[2], oa = F[3], ot = F[4];
            kfn rDw = pick_rD((int)P[0]);
            const double m = P[1], h = P[2], mu = P[3], c2 = P[4], rho0 = P[5];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                if (p.f[ot] == 0.0) {
                    double ker = m * rDw(h, r);
                    double Pp = c2 * (p.f[orho] - rho0), Pq = c2 * (q.f[orho] - rho0);
                    double a = -ker * (Pp / (p.f[orho] * p.f[orho]) + Pq / (q.f[orho] * q.f[orho]));
                    for (int c = 0; c < 3; c++) p.f[oa + c] += a * xpq[c];
                    double b = ker * 2 * mu / (p.f[orho] * q.f[orho]);
                    for (int c = 0; c < 3; c++) p.f[oa + c] += b * (p.f[ov + c] - q.f[ov + c]);
                }
            });
            return SP_OK;
        }
        case SP_OP_FIND_NORMAL: {  // drop.jl:76-78
            if (!need(2, 3)) return SP_ERR_INVALID;
            const int on = F[1];
            kfn rDw = pick_rD((int)P[0]);
            const double coef = P[1], h = P[2];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle&, const double* xpq, double r) {
                double k = coef * rDw(h, r);
                for (int c = 0; c < 3; c++) p.f[on + c] += k * xpq[c];
            });
            if (self) apply_unary(s, [=](Particle& p) { for (int c = 0; c < 3; c++) p.f[on + c] += (coef * rDw(h, 0.0)) * 0.0; });
            return SP_OK;
        }
        case SP_OP_NORMALIZE: {  // drop.jl:84-87
            if (!need(1, 1)) return SP_ERR_INVALID;
            const int on = F[0];
            const double s0 = P[0];
            apply_unary(s, [=](Particle& p) {
                double sn = std::sqrt(dot3(&p.f[on], &p.f[on]));
                for (int c = 0; c < 3; c++) p.f[on + c] /= (sn + s0);
            });
            return SP_OK;
        }
        case SP_OP_INTERNAL_FORCE_TENSION: {  // drop.jl:101-113
            if (!need(5, 6)) return SP_ERR_INVALID;
            const int ov = F[1], oP = F[2], on = F[3], oa = F[4];
            const double m = P[0], h = P[1], mu = P[2], rho0 = P[3], beta = P[4], s0 = P[5];
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double ker = m * rDwendland3(h, r);
                double a = -ker * (p.f[oP] / (rho0 * rho0) + q.f[oP] / (rho0 * rho0));
                for (int c = 0; c < 3; c++) p.f[oa + c] += a * xpq[c];
                double b = 2 * ker * mu / (rho0 * rho0);
                for (int c = 0; c < 3; c++) p.f[oa + c] += b * (p.f[ov + c] - q.f[ov + c]);
                double npq[3] = {p.f[on] - q.f[on], p.f[on + 1] - q.f[on + 1], p.f[on + 2] - q.f[on + 2]};
                double w = (m * DDwendland3(h, r) - ker) * dot3(xpq, npq);
                double t = 2 * beta / (rho0 * rho0);
                for (int c = 0; c < 3; c++) p.f[oa + c] -= t * (w * xpq[c] / (r * r + s0) + ker * npq[c]);
            });
            return SP_OK;
        }
        case SP_OP_MOVE_ALL: {  // static_container.jl:116-119
            if (!need(3, 1)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], oa = F[2];
            const double dtm = P[0];
            apply_unary(s, [=](Particle& p) {
                for (int c = 0; c < 3; c++) p.f[ox + c] += dtm * p.f[ov + c];
                p.f[oa] = p.f[oa + 1] = p.f[oa + 2] = 0.0;
            });
            return SP_OK;
        }
        case SP_OP_DENSITY_SUM_FLUID: {  // collapse_symplectic.jl:98-108, Kepler_vortex.jl:139-149
            if (!need(3, 3)) return SP_ERR_INVALID;
            const int oo = F[1], ot = F[2];
            kfn w = pick_w((int)P[0]);
            const double m = P[1], h = P[2];
            if (!w) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double*, double r) {
                if (p.f[ot] == 0.0 && q.f[ot] == 0.0) p.f[oo] += m * w(h, r);
            });
            if (self)  // action!(p, p, 0.0), core.jl:155-157
                apply_unary(s, [=](Particle& p) {
                    if (p.f[ot] == 0.0) p.f[oo] += m * w(h, 0.0);
                });
            return SP_OK;
        }
        case SP_OP_INTERNAL_FORCE_LJ: {  // collapse_symplectic.jl:114-123, Kepler_vortex.jl:155-164
            if (!need(5, 8)) return SP_ERR_INVALID;
            const int oP = F[1], orho = F[2], oa = F[3], ot = F[4];
            kfn rDw = pick_rD((int)P[0]);
            const double m = P[1], h = P[2], rho0 = P[3], wall = P[4], dr_wall = P[5], E_wall = P[6], eps = P[7];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                if (p.f[ot] == 0.0 && q.f[ot] == 0.0) {
                    double ker = m * rDw(h, r);
                    double a = (rho0 == 0.0)
                                   ? -ker * (p.f[oP] / (p.f[orho] * p.f[orho]) + q.f[oP] / (q.f[orho] * q.f[orho]))
                                   : -ker * (p.f[oP] / (rho0 * rho0) + q.f[oP] / (rho0 * rho0));
                    for (int c = 0; c < 3; c++) p.f[oa + c] += a * xpq[c];
                } else if (p.f[ot] == 0.0 && q.f[ot] == wall && r < dr_wall) {
                    double s_ = dr_wall / (r + eps);
                    double a = -E_wall / ((r + eps) * (r + eps)) * (s_ * s_ - julia_pow4(s_));
                    for (int c = 0; c < 3; c++) p.f[oa + c] += a * xpq[c];
                }
            });
            return SP_OK;
        }
        case SP_OP_MOVE_REV: {  // collapse_symplectic.jl:134-138, Kepler_vortex.jl:174-178
            if (!need(3, 1)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], ot = F[2];
            const double dt = P[0];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0)
                    for (int c = 0; c < 3; c++) p.f[ox + c] = rev_add(p.f[ox + c], dt * p.f[ov + c]);
            });
            return SP_OK;
        }
        case SP_OP_ACCELERATE_REV: {  // collapse_symplectic.jl:140-144
            if (!need(3, 4)) return SP_ERR_INVALID;
            const int ov = F[0], oa = F[1], ot = F[2];
            const double hdt = P[0], g[3] = {P[1], P[2], P[3]};
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0)
                    for (int c = 0; c < 3; c++) p.f[ov + c] = rev_add(p.f[ov + c], hdt * (p.f[oa + c] + g[c]));
            });
            return SP_OK;
        }
        case SP_OP_ACCELERATE_REV_CENTRAL: {  // Kepler_vortex.jl:180-184
            if (!need(4, 2)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], oa = F[2], ot = F[3];
            const double hdt = P[0], GM = P[1];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0) {
                    double n = std::sqrt(dot3(p.f + ox, p.f + ox));  // norm, algebra.jl:58-60
                    double k = -GM / (n * n * n);
                    for (int c = 0; c < 3; c++)
                        p.f[ov + c] = rev_add(p.f[ov + c], hdt * rev_add(p.f[oa + c], k * p.f[ox + c]));
                }
            });
            return SP_OK;
        }
        case SP_OP_LJ_POTENTIAL: {  // sum(sys, LJ_potential, p): core.jl:271-291 with collapse_symplectic.jl:146-153
            if (!need(3, 5)) return SP_ERR_INVALID;
            const int oo = F[1], ot = F[2];
            const double coef = P[1], wall = P[2], dr_wall = P[3], eps = P[4];
            // sum() does not skip q === p; LJ_potential(p, p, 0) = 0 because p cannot be both fluid and wall
            apply_binary(s, [=](Particle& p, const Particle& q, const double*, double r) {
                if (q.f[ot] == wall && p.f[ot] == 0.0 && r < dr_wall) {
                    double s_ = dr_wall / (r + eps);
                    p.f[oo] += coef * (0.5 * (s_ * s_) - 0.25 * julia_pow4(s_) - 0.25);
                }
            });
            return SP_OK;
        }
        case SP_OP_CYL_BALANCE_OF_MASS: {  // cylinder.jl:102-108
            if (!need(6, 3)) return SP_ERR_INVALID;
            const int ov = F[1], orho = F[2], oD = F[3], om = F[4], ot = F[5];
            kfn rDw = pick_rD((int)P[0]);
            const double h = P[1], two_nu = P[2];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double ker = q.f[om] * rDw(h, r);
                double vpq[3] = {p.f[ov] - q.f[ov], p.f[ov + 1] - q.f[ov + 1], p.f[ov + 2] - q.f[ov + 2]};
                p.f[oD] += ker * dot3(xpq, vpq);
                if (p.f[ot] == 0.0 && q.f[ot] == 0.0) p.f[oD] += two_nu / p.f[orho] * (p.f[orho] - q.f[orho]);
            });
            return SP_OK;
        }
        case SP_OP_CYL_FIND_PRESSURE: {  // cylinder.jl:110-116
            if (!need(4, 4)) return SP_ERR_INVALID;
            const int ox = F[0], orho = F[1], oD = F[2], oP = F[3];
            const double dt = P[0], c2 = P[1], rho0 = P[2], x1_min = P[3];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ox] >= x1_min) p.f[orho] += p.f[oD] * dt;
                p.f[oD] = 0.0;
                p.f[oP] = c2 * (p.f[orho] - rho0);
            });
            return SP_OK;
        }
        case SP_OP_CYL_INTERNAL_FORCE: {  // cylinder.jl:118-123
            if (!need(6, 4)) return SP_ERR_INVALID;
            const int ov = F[1], oP = F[2], orho = F[3], oa = F[4], om = F[5];
            kfn rDw = pick_rD((int)P[0]);
            const double h = P[1], mu = P[2], eps2 = P[3];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double ker = q.f[om] * rDw(h, r);
                double a = -ker * (p.f[oP] / (p.f[orho] * p.f[orho]) + q.f[oP] / (q.f[orho] * q.f[orho]));
                for (int c = 0; c < 3; c++) p.f[oa + c] += a * xpq[c];
                double vpq[3] = {p.f[ov] - q.f[ov], p.f[ov + 1] - q.f[ov + 1], p.f[ov + 2] - q.f[ov + 2]};
                double b = 8.0 * ker * mu / (p.f[orho] * q.f[orho]) * dot3(vpq, xpq) / (r * r + eps2);
                for (int c = 0; c < 3; c++) p.f[oa + c] += b * xpq[c];
            });
            return SP_OK;
        }
        case SP_OP_MOVE_TYPES: {  // cylinder.jl:125-130
            if (!need(4, 3)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], oa = F[2], ot = F[3];
            const double dt = P[0], ta = P[1], tb = P[2];
            apply_unary(s, [=](Particle& p) {
                p.f[oa] = p.f[oa + 1] = p.f[oa + 2] = 0.0;
                if (p.f[ot] == ta || p.f[ot] == tb)
                    for (int c = 0; c < 3; c++) p.f[ox + c] += dt * p.f[ov + c];
            });
            return SP_OK;
        }
        case SP_OP_CYL_ACCELERATE: {  // cylinder.jl:132-143 (gravity: :132-137)
            if (!need(4, 3)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], oa = F[2], ot = F[3];
            const double hdt = P[0], cyl1 = P[1], coef = P[2];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0) {
                    double f[3] = {cyl1 - p.f[ox], -p.f[ox + 1], 0.0};
                    double absf2 = (cyl1 - p.f[ox]) * (cyl1 - p.f[ox]) + p.f[ox + 1] * p.f[ox + 1];
                    for (int c = 0; c < 3; c++) p.f[ov + c] += hdt * (p.f[oa + c] + coef * f[c] / absf2);
                }
            });
            return SP_OK;
        }
        case SP_OP_SET_INFLOW_SPEED: {  // cylinder.jl:91-97
            if (!need(3, 4)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], ot = F[2];
            const double inflow = P[0], sfac = P[1], U_max = P[2], chan_w = P[3];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == inflow) {
                    double q = 2.0 * p.f[ox + 1] / chan_w;
                    double v1 = sfac * U_max * (1.0 - q * q);
                    p.f[ov] = v1 * 1.0;
                    p.f[ov + 1] = v1 * 0.0;
                    p.f[ov + 2] = v1 * 0.0;
                }
            });
            return SP_OK;
        }
        case SP_OP_ROD_FIND_A: {  // rod.jl:128-134
            if (!need(4, 2)) return SP_ERR_INVALID;
            const int oX = F[1], oA = F[2], oH = F[3];
            kfn w = pick_w((int)P[0]);
            const double h = P[1];
            if (!w) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double ker = w(h, r);
                double Xpq[2] = {p.f[oX] - q.f[oX], p.f[oX + 1] - q.f[oX + 1]};
                // outer(x, y)[i,j] = x[i]*y[j]  (:44-50)
                p.f[oA + 0] += -ker * (Xpq[0] * xpq[0]);
                p.f[oA + 1] += -ker * (Xpq[1] * xpq[0]);
                p.f[oA + 3] += -ker * (Xpq[0] * xpq[1]);
                p.f[oA + 4] += -ker * (Xpq[1] * xpq[1]);
                p.f[oH + 0] += -ker * (xpq[0] * xpq[0]);
                p.f[oH + 1] += -ker * (xpq[1] * xpq[0]);
                p.f[oH + 3] += -ker * (xpq[0] * xpq[1]);
                p.f[oH + 4] += -ker * (xpq[1] * xpq[1]);
            });
            return SP_OK;
        }
        case SP_OP_ROD_FIND_B: {  // rod.jl:136-143
            if (!need(3, 3)) return SP_ERR_INVALID;
            const int oA = F[0], oH = F[1], oB = F[2];
            const double m = P[0], c_l = P[1], c_s = P[2];
            apply_unary(s, [=](Particle& p) {
                M2 Hi = m2_inv(m2_load(p.f + oH));
                M2 A = m2_mul(m2_load(p.f + oA), Hi);
                m2_store(p.f + oA, A);
                M2 At = m2_trans(A);
                M2 G = m2_mul(At, A);
                double Pr = (c_l * c_l) * (m2_det(A) - 1.0);
                M2 B = m2_mul(m2_scale(m, m2_add(m2_scale(Pr, m2_inv(At)), m2_mul(m2_scale(c_s * c_s, A), m2_dev(G, nullptr)))), Hi);
                m2_store(p.f + oB, B);
            });
            return SP_OK;
        }
        case SP_OP_ROD_FIND_F: {  // rod.jl:145-160
            if (!need(6, 4)) return SP_ERR_INVALID;
            const int ov = F[1], oX = F[2], oA = F[3], oB = F[4], of = F[5];
            kfn w = pick_w((int)P[0]), rDw = pick_rD((int)P[0]);
            const double h = P[1], two_m_vol = P[2], nu = P[3];
            if (!w || !rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double ker = w(h, r), rDker = rDw(h, r);
                double Xpq[2] = {p.f[oX] - q.f[oX], p.f[oX + 1] - q.f[oX + 1]};
                M2 Ap = m2_load(p.f + oA), Bp = m2_load(p.f + oB), Aq = m2_load(q.f + oA), Bq = m2_load(q.f + oB);
                double y[2], z[2];
                m2_vec(Bp, xpq, y);
                m2_vec(m2_trans(Ap), y, z);
                p.f[of] += -ker * z[0];
                p.f[of + 1] += -ker * z[1];
                m2_vec(Bq, xpq, y);
                m2_vec(m2_trans(Aq), y, z);
                p.f[of] += -ker * z[0];
                p.f[of + 1] += -ker * z[1];
                // "eta" correction
                double ax[2], wv[2], kpq[2], kqp[2];
                m2_vec(Ap, xpq, ax);
                wv[0] = Xpq[0] - ax[0]; wv[1] = Xpq[1] - ax[1];
                m2_vec(m2_trans(Bp), wv, kpq);
                m2_vec(Aq, xpq, ax);
                wv[0] = Xpq[0] - ax[0]; wv[1] = Xpq[1] - ax[1];
                m2_vec(m2_trans(Bq), wv, kqp);
                kqp[0] = -kqp[0]; kqp[1] = -kqp[1];
                double dp = xpq[0] * kpq[0] + xpq[1] * kpq[1], dq = xpq[0] * kqp[0] + xpq[1] * kqp[1];
                for (int c = 0; c < 2; c++) p.f[of + c] += rDker * dp * xpq[c] + ker * kpq[c];
                for (int c = 0; c < 2; c++) p.f[of + c] -= rDker * dq * xpq[c] + ker * kqp[c];
                // artificial viscosity
                double visc = two_m_vol * rDker * nu;
                for (int c = 0; c < 3; c++) p.f[of + c] += visc * (p.f[ov + c] - q.f[ov + c]);
            });
            return SP_OK;
        }
        case SP_OP_ROD_PULL: {  // rod.jl:162-166
            if (!need(2, 2)) return SP_ERR_INVALID;
            const int oX = F[0], of = F[1];
            const double X1_min = P[0], fy = P[1];
            apply_unary(s, [=](Particle& p) {
                if (p.f[oX] > X1_min) p.f[of + 1] += fy;
            });
            return SP_OK;
        }
        case SP_OP_ROD_UPDATE_V: {  // rod.jl:168-174
            if (!need(3, 3)) return SP_ERR_INVALID;
            const int ov = F[0], of = F[1], oX = F[2];
            const double hdt = P[0], m = P[1], X1_clamp = P[2];
            apply_unary(s, [=](Particle& p) {
                for (int c = 0; c < 3; c++) p.f[ov + c] += hdt * p.f[of + c] / m;
                if (p.f[oX] < X1_clamp) p.f[ov] = p.f[ov + 1] = p.f[ov + 2] = 0.0;
            });
            return SP_OK;
        }
        case SP_OP_ROD_UPDATE_X: {  // rod.jl:176-183
            if (!need(6, 1)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], oA = F[2], oH = F[3], of = F[4], oe = F[5];
            const double dt = P[0];
            apply_unary(s, [=](Particle& p) {
                for (int c = 0; c < 3; c++) p.f[ox + c] += dt * p.f[ov + c];
                for (int c = 0; c < 9; c++) p.f[oH + c] = p.f[oA + c] = 0.0;
                p.f[of] = p.f[of + 1] = p.f[of + 2] = 0.0;
                p.f[oe] = 0.0;
            });
            return SP_OK;
        }
        case SP_OP_ROD_FIND_E: {  // rod.jl:185-188
            if (!need(4, 1)) return SP_ERR_INVALID;
            const int oX = F[1], oA = F[2], oe = F[3];
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double) {
                double Xpq[2] = {p.f[oX] - q.f[oX], p.f[oX + 1] - q.f[oX + 1]}, y[2];
                m2_vec(m2_inv(m2_load(p.f + oA)), Xpq, y);
                double eta[3] = {y[0] - xpq[0], y[1] - xpq[1], 0.0 - xpq[2]};
                p.f[oe] += dot3(eta, eta);
            });
            return SP_OK;
        }
        case SP_OP_SHTC_FIND_STRESS: {  // SHTC/ldc.jl:118-121
            if (!need(3, 3)) return SP_ERR_INVALID;
            const int oA = F[0], orho = F[1], oS = F[2];
            const double c_l = P[0], c_s = P[1], rho_ref = P[2];
            apply_unary(s, [=](Particle& p) {
                M3 finger = m3_tmul(m3_load(p.f + oA), m3_load(p.f + oA));
                M3 S = m3_mul(m3_scale((c_s * c_s) * p.f[orho], finger), m3_dev(finger));
                const double iso = (c_l * c_l) * (p.f[orho] - rho_ref);
                S.a[0] = iso + S.a[0];
                S.a[4] = iso + S.a[4];
                S.a[8] = iso + S.a[8];
                m3_store(p.f + oS, S);
            });
            return SP_OK;
        }
        case SP_OP_SHTC_UPDATE_V: {  // SHTC/ldc.jl:123-127
            if (!need(5, 3)) return SP_ERR_INVALID;
            const int ov = F[1], orho = F[2], oS = F[3], ot = F[4];
            kfn rDw = pick_rD((int)P[0]);
            const double h = P[1], dtm = P[2];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                if (p.f[ot] == 0.0) {
                    const double c = -dtm * rDw(h, r);
                    const double rp2 = p.f[orho] * p.f[orho], rq2 = q.f[orho] * q.f[orho];
                    M3 S;
                    for (int k = 0; k < 9; k++) S.a[k] = c * (p.f[oS + k] / rp2 + q.f[oS + k] / rq2);
                    for (int i = 0; i < 3; i++) p.f[ov + i] += S.a[i] * xpq[0] + S.a[i + 3] * xpq[1] + S.a[i + 6] * xpq[2];
                }
            });
            return SP_OK;
        }
        case SP_OP_SHTC_UPDATE_RHO: {  // SHTC/ldc.jl:90-94
            if (!need(4, 3)) return SP_ERR_INVALID;
            const int ov = F[1], orho = F[2], ot = F[3];
            kfn rDw = pick_rD((int)P[0]);
            const double h = P[1], dtm = P[2];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                if (p.f[ot] == 0.0) {
                    double vpq[3] = {p.f[ov] - q.f[ov], p.f[ov + 1] - q.f[ov + 1], p.f[ov + 2] - q.f[ov + 2]};
                    p.f[orho] += dtm * rDw(h, r) * dot3(xpq, vpq);
                }
            });
            return SP_OK;
        }
        case SP_OP_SHTC_CONVECT_A: {  // SHTC/ldc.jl:96-100 (uses the A_p left by the previous pairs: order-dependent)
            if (!need(5, 4)) return SP_ERR_INVALID;
            const int ov = F[1], orho = F[2], oA = F[3], ot = F[4];
            kfn rDw = pick_rD((int)P[0]);
            const double h = P[1], dtm = P[2], skip = P[3];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                if (p.f[ot] != skip) {
                    M3 A = m3_load(p.f + oA), M;
                    for (int j = 0; j < 3; j++)
                        for (int i = 0; i < 3; i++) M.a[i + 3 * j] = (p.f[ov + i] - q.f[ov + i]) * xpq[j];  // v_pq*x_pq'
                    m3_store(p.f + oA, m3_add(A, m3_mul(m3_scale(dtm / p.f[orho] * rDw(h, r), A), M)));
                }
            });
            return SP_OK;
        }
        case SP_OP_SHTC_RELAX_A: {  // SHTC/ldc.jl:106-116 (RK4)
            if (!need(1, 2)) return SP_ERR_INVALID;
            const int oA = F[0];
            const double dt = P[0], tau = P[1];
            apply_unary(s, [=](Particle& p) {
                const M3 A = m3_load(p.f + oA);
                M3 K;
                // A_new += dt*K/6 ; K = f(A + dt*K/2) ; A_new += dt*K/3 ; K = f(A + dt*K/2) ; A_new += dt*K/3 ;
                // K = f(A + dt*K) ; A_new += dt*K/6          (dt*K/n is (dt*K)/n, element by element)
                auto over = [](const M3& X, double n) {
                    M3 C;
                    for (int k = 0; k < 9; k++) C.a[k] = X.a[k] / n;
                    return C;
                };
                M3 R = A;
                K = shtc_relax_f(A, tau);
                R = m3_add(R, over(m3_scale(dt, K), 6.0));
                K = shtc_relax_f(m3_add(A, over(m3_scale(dt, K), 2.0)), tau);
                R = m3_add(R, over(m3_scale(dt, K), 3.0));
                K = shtc_relax_f(m3_add(A, over(m3_scale(dt, K), 2.0)), tau);
                R = m3_add(R, over(m3_scale(dt, K), 3.0));
                K = shtc_relax_f(m3_add(A, m3_scale(dt, K)), tau);
                R = m3_add(R, over(m3_scale(dt, K), 6.0));
                m3_store(p.f + oA, R);
            });
            return SP_OK;
        }
        case SP_OP_SHTC_MOVE: {  // SHTC/ldc.jl:129-133
            if (!need(3, 1)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], ot = F[2];
            const double dt = P[0];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0)
                    for (int c = 0; c < 3; c++) p.f[ox + c] += p.f[ov + c] * dt;
            });
            return SP_OK;
        }
        case SP_OP_BE_FIND_L: {  // SHTC/beryllium.jl:140-146
            if (!need(5, 3)) return SP_ERR_INVALID;
            const int ov = F[1], om = F[2], oT = F[3], oL = F[4];
            kfn rDw = pick_rD((int)P[0]);
            const double h = P[1], rho0 = P[2];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                const double ker = q.f[om] / rho0 * rDw(h, r);
                const double vpq[2] = {p.f[ov] - q.f[ov], p.f[ov + 1] - q.f[ov + 1]};
                // outer(x, y)[i,j] = x[i]*y[j], in-plane block only (:79-85)
                p.f[oT + 0] += ker * (xpq[0] * xpq[0]);
                p.f[oT + 1] += ker * (xpq[1] * xpq[0]);
                p.f[oT + 3] += ker * (xpq[0] * xpq[1]);
                p.f[oT + 4] += ker * (xpq[1] * xpq[1]);
                p.f[oL + 0] += ker * (vpq[0] * xpq[0]);
                p.f[oL + 1] += ker * (vpq[1] * xpq[0]);
                p.f[oL + 3] += ker * (vpq[0] * xpq[1]);
                p.f[oL + 4] += ker * (vpq[1] * xpq[1]);
            });
            return SP_OK;
        }
        case SP_OP_BE_UPDATE_A: {  // SHTC/beryllium.jl:148-151
            if (!need(3, 1)) return SP_ERR_INVALID;
            const int oA = F[0], oT = F[1], oL = F[2];
            const double hdt = P[0];
            apply_unary(s, [=](Particle& p) {
                // L = L*inv(T): both in-plane (inv's [3,3] = 1 meets a zero column of L)
                const M2 L = m2_mul(m2_load(p.f + oL), m2_inv(m2_load(p.f + oT)));
                m2_store(p.f + oL, L);
                // A = A*(MAT1 - hdt*L)*inv(MAT1 + hdt*L): block diagonal, the [3,3] entry is A33*1*1
                const M2 I2{1.0, 0.0, 0.0, 1.0};
                const M2 hL = m2_scale(hdt, L);
                const M2 minus{I2.a11 - hL.a11, I2.a21 - hL.a21, I2.a12 - hL.a12, I2.a22 - hL.a22};
                const M2 A = m2_mul(m2_mul(m2_load(p.f + oA), minus), m2_inv(m2_add(I2, hL)));
                const double a33 = p.f[oA + 8];
                m2_store(p.f + oA, A);
                p.f[oA + 8] = a33;
            });
            return SP_OK;
        }
        case SP_OP_BE_FIND_J: {  // SHTC/beryllium.jl:153-158
            if (!need(5, 3)) return SP_ERR_INVALID;
            const int om = F[1], oT = F[2], oJ = F[3], oK = F[4];
            kfn rDw = pick_rD((int)P[0]), w = pick_w((int)P[0]);
            const double h = P[1], rho0 = P[2];
            if (!rDw || !w) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                const double ker = q.f[om] / rho0 * rDw(h, r);
                p.f[oT + 0] += ker * (xpq[0] * xpq[0]);
                p.f[oT + 1] += ker * (xpq[1] * xpq[0]);
                p.f[oT + 3] += ker * (xpq[0] * xpq[1]);
                p.f[oT + 4] += ker * (xpq[1] * xpq[1]);
                p.f[oJ] += q.f[om] / rho0 * w(h, r);
                p.f[oK] += q.f[om] / rho0 * wendland2h(h, r);
            });
            if (self)  // taco.jl:252 find_rho!(p, p, 0.0): x_pq = 0 adds nothing to T
                apply_unary(s, [=](Particle& p) {
                    p.f[oJ] += p.f[om] / rho0 * w(h, 0.0);
                    p.f[oK] += p.f[om] / rho0 * wendland2h(h, 0.0);
                });
            return SP_OK;
        }
        case SP_OP_BE_FIND_T: {  // SHTC/beryllium.jl:160-164
            if (!need(4, 3)) return SP_ERR_INVALID;
            const int oA = F[0], oT = F[1], oP = F[2], oJ = F[3];
            const double rho0 = P[0], c_0 = P[1], c_s = P[2];
            apply_unary(s, [=](Particle& p) {
                const M2 A = m2_load(p.f + oA);
                const double a33 = p.f[oA + 8];
                const M2 G = m2_mul(m2_trans(A), A);
                const double g33 = a33 * a33;
                const double J = p.f[oJ];
                const double Pr = 0.5 * rho0 * (c_0 * c_0) * ((1.0 - 1.0 / J) / (J * J) + std::log(J) / J);
                p.f[oP] = Pr;
                const double tr = 1.0 / 3.0 * (G.a11 + G.a22 + g33);  // dev :100-103
                const M2 D{G.a11 - tr, G.a21, G.a12, G.a22 - tr};
                const M2 S = m2_mul(m2_mul(m2_scale(c_s * c_s, G), D), m2_inv(m2_load(p.f + oT)));
                const double s33 = (c_s * c_s) * g33 * (g33 - tr) * 1.0;  // inv(T)[3,3] = 1 (:91-98)
                const double iso = Pr / rho0;
                m2_store(p.f + oT, M2{iso - S.a11, -S.a21, -S.a12, iso - S.a22});
                p.f[oT + 8] = iso - s33;
            });
            return SP_OK;
        }
        case SP_OP_BE_FIND_F: {  // SHTC/beryllium.jl:166-175
            if (!need(5, 4)) return SP_ERR_INVALID;
            const int om = F[1], oT = F[2], oK = F[3], of = F[4];
            kfn rDw = pick_rD((int)P[0]);
            const double h = P[1], rho0 = P[2], c_p = P[3];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                const double ker = q.f[om] / rho0 * rDw(h, r);
                const double kerh = q.f[om] / rho0 * rDwendland2h(h, r);
                double y[2];
                m2_vec(m2_load(p.f + oT), xpq, y);
                p.f[of] += -p.f[om] * ker * y[0];
                p.f[of + 1] += -p.f[om] * ker * y[1];
                m2_vec(m2_load(q.f + oT), xpq, y);
                p.f[of] += -p.f[om] * ker * y[0];
                p.f[of + 1] += -p.f[om] * ker * y[1];
                const double a = -p.f[om] * kerh * (c_p * c_p) * (p.f[oK] + q.f[oK]);
                for (int c = 0; c < 3; c++) p.f[of + c] += a * xpq[c];
            });
            return SP_OK;
        }
        case SP_OP_BE_RESET: {  // SHTC/beryllium.jl:177-184
            if (!need(7, 0)) return SP_ERR_INVALID;
            const int of = F[0], oL = F[1], oT = F[2], oJ = F[3], oK = F[4], oJ0 = F[5], oK0 = F[6];
            apply_unary(s, [=](Particle& p) {
                p.f[of] = p.f[of + 1] = p.f[of + 2] = 0.0;
                for (int c = 0; c < 9; c++) p.f[oL + c] = p.f[oT + c] = 0.0;
                p.f[oJ] = p.f[oJ0];
                p.f[oK] = p.f[oK0];
            });
            return SP_OK;
        }
        case SP_OP_BE_UPDATE_V: {  // SHTC/beryllium.jl:132-134
            if (!need(3, 1)) return SP_ERR_INVALID;
            const int ov = F[0], of = F[1], om = F[2];
            const double hdt = P[0];
            apply_unary(s, [=](Particle& p) {
                for (int c = 0; c < 3; c++) p.f[ov + c] += hdt * p.f[of + c] / p.f[om];
            });
            return SP_OK;
        }
        case SP_OP_TW_FIND_L:    // SHTC/twist3d.jl:135-141
        case SP_OP_TW_FIND_J: {  // SHTC/twist3d.jl:148-153
            if (!need(5, 3)) return SP_ERR_INVALID;
            const bool withL = op == SP_OP_TW_FIND_L;
            const int ov = withL ? F[1] : 0, om = withL ? F[2] : F[1], oT = withL ? F[3] : F[2];
            const int oL = withL ? F[4] : 0, oJ = withL ? 0 : F[3], oK = withL ? 0 : F[4];
            kfn rDw = pick_rD((int)P[0]), w = pick_w((int)P[0]);
            const double h = P[1], rho0 = P[2];
            if (!rDw || !w) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                const double ker = q.f[om] / rho0 * rDw(h, r);
                for (int j = 0; j < 3; j++)
                    for (int i = 0; i < 3; i++) p.f[oT + i + 3 * j] += ker * (xpq[i] * xpq[j]);  // outer(x, y)[i,j] = x[i]*y[j]
                if (withL) {
                    for (int j = 0; j < 3; j++)
                        for (int i = 0; i < 3; i++) p.f[oL + i + 3 * j] += ker * ((p.f[ov + i] - q.f[ov + i]) * xpq[j]);
                } else {
                    p.f[oJ] += q.f[om] / rho0 * w(h, r);
                    p.f[oK] += q.f[om] / rho0 * wendland3h(h, r);
                }
            });
            return SP_OK;
        }
        case SP_OP_TW_UPDATE_A: {  // SHTC/twist3d.jl:143-146
            if (!need(3, 1)) return SP_ERR_INVALID;
            const int oA = F[0], oT = F[1], oL = F[2];
            const double hdt = P[0];
            apply_unary(s, [=](Particle& p) {
                const M3 L = m3_mul(m3_load(p.f + oL), m3_inv(m3_load(p.f + oT)));
                m3_store(p.f + oL, L);
                const M3 I = m3_identity(), hL = m3_scale(hdt, L);
                M3 minus;
                for (int k = 0; k < 9; k++) minus.a[k] = I.a[k] - hL.a[k];
                m3_store(p.f + oA, m3_mul(m3_mul(m3_load(p.f + oA), minus), m3_inv(m3_add(I, hL))));
            });
            return SP_OK;
        }
        case SP_OP_TW_FIND_T: {  // SHTC/twist3d.jl:155-161
            if (!need(4, 3)) return SP_ERR_INVALID;
            const int oA = F[0], oT = F[1], oP = F[2], oJ = F[3];
            const double rho0 = P[0], c_0 = P[1], c_s = P[2];
            apply_unary(s, [=](Particle& p) {
                const M3 Fm = m3_inv(m3_load(p.f + oA));
                M3 B;  // F*F'
                for (int j = 0; j < 3; j++)
                    for (int i = 0; i < 3; i++)
                        B.a[i + 3 * j] = Fm.a[i] * Fm.a[j] + Fm.a[i + 3] * Fm.a[j + 3] + Fm.a[i + 6] * Fm.a[j + 6];
                const double detF = 1.0 / p.f[oJ];
                const double Pr = -rho0 * (c_0 * c_0) * (detF * detF) * (detF - 1.0);
                p.f[oP] = Pr;
                const M3 I = m3_identity();
                M3 BmI;
                for (int k = 0; k < 9; k++) BmI.a[k] = B.a[k] - I.a[k];
                const M3 S = m3_mul(m3_scale(c_s * c_s, BmI), m3_inv(m3_load(p.f + oT)));
                M3 T;
                for (int k = 0; k < 9; k++) T.a[k] = -Pr / rho0 * I.a[k] - S.a[k];
                m3_store(p.f + oT, T);
            });
            return SP_OK;
        }
        case SP_OP_TW_FIND_F: {  // SHTC/twist3d.jl:163-172
            if (!need(5, 4)) return SP_ERR_INVALID;
            const int om = F[1], oT = F[2], oK = F[3], of = F[4];
            kfn rDw = pick_rD((int)P[0]);
            const double h = P[1], rho0 = P[2], c_p = P[3];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                const double ker = q.f[om] / rho0 * rDw(h, r);
                const double kerh = q.f[om] / rho0 * rDwendland3h(h, r);
                for (int pass = 0; pass < 2; pass++) {
                    const double* T = (pass == 0 ? p.f : q.f) + oT;
                    for (int i = 0; i < 3; i++)
                        p.f[of + i] += p.f[om] * ker * (T[i] * xpq[0] + T[i + 3] * xpq[1] + T[i + 6] * xpq[2]);
                }
                const double a = -p.f[om] * kerh * (c_p * c_p) * (p.f[oK] + q.f[oK]);
                for (int c = 0; c < 3; c++) p.f[of + c] += a * xpq[c];
            });
            return SP_OK;
        }
        case SP_OP_TW_UPDATE_V: {  // SHTC/twist3d.jl:125-129
            if (!need(4, 1)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], of = F[2], om = F[3];
            const double hdt = P[0];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ox + 2] > 0.0)
                    for (int c = 0; c < 3; c++) p.f[ov + c] += hdt * p.f[of + c] / p.f[om];
            });
            return SP_OK;
        }
        case SP_OP_TA_FIND_T: {  // SHTC/taco.jl:148-152 (subinv: tools.jl:45-52, zeros outside the in-plane block)
            if (!need(4, 3)) return SP_ERR_INVALID;
            const int oA = F[0], oT = F[1], oP = F[2], orho = F[3];
            const double rho0 = P[0], c_0 = P[1], c_s = P[2];
            apply_unary(s, [=](Particle& p) {
                const M3 A = m3_load(p.f + oA);
                const M3 G = m3_tmul(A, A);
                const M3 GD = m3_mul(m3_scale(c_s * c_s, G), m3_dev(G));
                const double rho = p.f[orho];
                const double Pr = (c_0 * c_0) * (rho - rho0) * rho0 / rho;
                p.f[oP] = Pr;
                const M2 si = m2_inv(m2_load(p.f + oT));
                M3 S;  // (c_s^2*G*dev(G))*subinv(T): the third column of subinv is zero
                for (int i = 0; i < 3; i++) {
                    S.a[i] = GD.a[i] * si.a11 + GD.a[i + 3] * si.a21;
                    S.a[i + 3] = GD.a[i] * si.a12 + GD.a[i + 3] * si.a22;
                    S.a[i + 6] = 0.0;
                }
                const double iso = -Pr / (rho * rho);
                S.a[0] = iso + S.a[0];
                S.a[4] = iso + S.a[4];
                S.a[8] = iso + S.a[8];
                m3_store(p.f + oT, S);
            });
            return SP_OK;
        }
        case SP_OP_TA_FIND_F: {  // SHTC/taco.jl:154-162
            if (!need(5, 3)) return SP_ERR_INVALID;
            const int om = F[1], oT = F[2], ol = F[3], of = F[4];
            kfn rDw = pick_rD((int)P[0]);
            const double h = P[1], cpr2 = P[2];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                const double ker = q.f[om] * rDw(h, r);
                const double kerh = q.f[om] * rDwendland2h(h, r);
                const double c = p.f[om] * ker;
                for (int i = 0; i < 3; i++) {
                    double row = 0.0;
                    for (int j = 0; j < 3; j++) row += (c * (p.f[oT + i + 3 * j] + q.f[oT + i + 3 * j])) * xpq[j];
                    p.f[of + i] += row;
                }
                const double a = -p.f[om] * kerh * cpr2 * (p.f[ol] + q.f[ol]);
                for (int i = 0; i < 3; i++) p.f[of + i] += a * xpq[i];
            });
            return SP_OK;
        }
        case SP_OP_TA_UPDATE_V: {  // SHTC/taco.jl:108-114 with vexact :39-42
            if (!need(5, 4)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], of = F[2], om = F[3], ot = F[4];
            const double hdt = P[0], R1 = P[1], R2 = P[2], omega = P[3];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0) {
                    for (int c = 0; c < 3; c++) p.f[ov + c] += hdt * p.f[of + c] / p.f[om];
                } else {
                    const double r = std::sqrt(dot3(p.f + ox, p.f + ox));
                    const double sc = R2 / r * (r / R1 - R1 / r) / (R2 / R1 - R1 / R2);
                    p.f[ov] = sc * (-omega * p.f[ox + 1]);
                    p.f[ov + 1] = sc * (omega * p.f[ox]);
                    p.f[ov + 2] = sc * 0.0;
                }
            });
            return SP_OK;
        }
        case SP_OP_TA_UPDATE_X: {  // SHTC/taco.jl:116-126
            if (!need(4, 4)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], ox0 = F[2], ot = F[3];
            const double hdt = P[0], cw = P[1], sw = P[2], outer = P[3];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0) {
                    for (int c = 0; c < 3; c++) p.f[ox + c] += hdt * p.f[ov + c];
                } else if (p.f[ot] == outer) {
                    const double a = p.f[ox0], b = p.f[ox0 + 1];
                    p.f[ox] = a * cw - b * sw;
                    p.f[ox + 1] = a * sw + b * cw;
                    p.f[ox + 2] = 0.0;
                }
            });
            return SP_OK;
        }
        case SP_OP_DENSITY_SUM: {  // test_collision_2d.jl:63-69; self term added last (core.jl:155-157)
            if (!need(2, 3)) return SP_ERR_INVALID;
            const int oo = F[1];
            kfn w = pick_w((int)P[0]);
            const double m = P[1], h = P[2];
            if (!w) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle&, const double*, double r) { p.f[oo] += m * w(h, r); });
            if (self) apply_unary(s, [=](Particle& p) { p.f[oo] += m * w(h, 0.0); });
            return SP_OK;
        }
        case SP_OP_PRESSURE_FROM_RHO: {  // test_collision_2d.jl:71-73
            if (!need(3, 1)) return SP_ERR_INVALID;
            const int orho = F[0], orho0 = F[1], oP = F[2];
            const double c2 = P[0];
            apply_unary(s, [=](Particle& p) { p.f[oP] = c2 * (p.f[orho] - p.f[orho0]); });
            return SP_OK;
        }
        case SP_OP_INTERNAL_FORCE_SYM: {  // test_collision_2d.jl:75-78
            if (!need(3, 4)) return SP_ERR_INVALID;
            const int oP = F[1], oa = F[2];
            kfn rDw = pick_rD((int)P[0]);
            const double m = P[1], h = P[2], rho0 = P[3];
            if (!rDw) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double ker = m * rDw(h, r);
                double a = -ker * (p.f[oP] / (rho0 * rho0) + q.f[oP] / (rho0 * rho0));
                for (int c = 0; c < 3; c++) p.f[oa + c] += a * xpq[c];
            });
            return SP_OK;
        }
        case SP_OP_FILL: {  // test_collision_2d.jl:80-86; F[1] = ncomp passed as second "field"
            if (!need(2, 1)) return SP_ERR_INVALID;
            const int of = F[0], nc = F[1];
            const double v = P[0];
            apply_unary(s, [=](Particle& p) {
                for (int c = 0; c < nc; c++) p.f[of + c] = v;
            });
            return SP_OK;
        }
        case SP_OP_ADVECT: {  // test_collision_2d.jl:88-90
            if (!need(2, 1)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1];
            const double dt = P[0];
            apply_unary(s, [=](Particle& p) {
                for (int c = 0; c < 3; c++) p.f[ox + c] += dt * p.f[ov + c];
            });
            return SP_OK;
        }
        case SP_OP_KICK: {  // test_collision_2d.jl:92-94
            if (!need(2, 1)) return SP_ERR_INVALID;
            const int ov = F[0], oa = F[1];
            const double hdt = P[0];
            apply_unary(s, [=](Particle& p) {
                for (int c = 0; c < 3; c++) p.f[ov + c] += hdt * p.f[oa + c];
            });
            return SP_OK;
        }
        case SP_OP_ISPH_INITIALIZE: {  // collapse_dry_implicit.jl:118-126
            if (!need(6, 4)) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], odiv = F[2], oL = F[3], olam = F[4], ot = F[5];
            const double dt = P[0], g[3] = {P[1], P[2], P[3]};
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0) {
                    for (int c = 0; c < 3; c++) p.f[ox + c] += dt * p.f[ov + c];
                    for (int c = 0; c < 3; c++) p.f[ov + c] += dt * g[c];
                }
                p.f[odiv] = 0.0;
                p.f[oL] = 0.0;
                p.f[olam] = 1.0;
            });
            return SP_OK;
        }
        case SP_OP_ISPH_VISCOUS_FORCE: {  // collapse_dry_implicit.jl:128-130
            if (!need(3, 5)) return SP_ERR_INVALID;
            const int ov = F[1], oDv = F[2];
            kfn rDk = pick_rD((int)P[0]);
            const double m = P[1], h = P[2], mu = P[3], rho = P[4];
            if (!rDk) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double*, double r) {
                double a = 2.0 * m * mu * rDk(h, r) / (rho * rho);
                for (int c = 0; c < 3; c++) p.f[oDv + c] += a * (p.f[ov + c] - q.f[ov + c]);
            });
            return SP_OK;
        }
        case SP_OP_ISPH_DIV_L_LAMBDA: {  // collapse_dry_implicit.jl:147-152
            if (!need(5, 5)) return SP_ERR_INVALID;
            const int ov = F[1], odiv = F[2], oL = F[3], olam = F[4];
            kfn rDkf = pick_rD((int)P[0]);
            const double m = P[1], h = P[2], rho = P[3], dim = P[4];
            if (!rDkf) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double rDk = rDkf(h, r);
                double vpq[3] = {p.f[ov] - q.f[ov], p.f[ov + 1] - q.f[ov + 1], p.f[ov + 2] - q.f[ov + 2]};
                p.f[odiv] += -dot3(xpq, vpq) * m * rDk;
                p.f[oL] += -2.0 * m / rho * rDk;
                p.f[olam] += m / rho * rDk * (r * r) / dim;
            });
            return SP_OK;
        }
        case SP_OP_ISPH_PROJECTION_VECTOR: {  // collapse_dry_implicit.jl:165-167 via assemble_vector core.jl:175-182
            if (!need(2, 2)) return SP_ERR_INVALID;
            const int odiv = F[0], ob = F[1];
            const double h = P[0], dt = P[1];
            apply_unary(s, [=](Particle& p) { p.f[ob] = -(h * h) * p.f[odiv] / dt; });
            return SP_OK;
        }
        case SP_OP_ISPH_INTERNAL_FORCE: {  // collapse_dry_implicit.jl:132-134
            if (!need(3, 4)) return SP_ERR_INVALID;
            const int oP = F[1], oDv = F[2];
            kfn rDk = pick_rD((int)P[0]);
            const double m = P[1], h = P[2], rho = P[3];
            if (!rDk) return SP_ERR_INVALID;
            apply_binary(s, [=](Particle& p, const Particle& q, const double* xpq, double r) {
                double a = m * rDk(h, r) * (p.f[oP] + q.f[oP]) / (rho * rho);
                for (int c = 0; c < 3; c++) p.f[oDv + c] -= a * xpq[c];
            });
            return SP_OK;
        }
        case SP_OP_ISPH_ACCELERATE: {  // collapse_dry_implicit.jl:136-141
            if (!need(3, 1)) return SP_ERR_INVALID;
            const int ov = F[0], oDv = F[1], ot = F[2];
            const double dt = P[0];
            apply_unary(s, [=](Particle& p) {
                if (p.f[ot] == 0.0)
                    for (int c = 0; c < 3; c++) p.f[ov + c] += dt * p.f[oDv + c];
                p.f[oDv] = p.f[oDv + 1] = p.f[oDv + 2] = 0.0;
            });
            return SP_OK;
        }
    }
    return SP_ERR_INVALID;
}

// projection_matrix, collapse_dry_implicit.jl:154-163
inline double projection_matrix(const Particle& p, bool same, double r, int oL, int olam, int ot, kfn rDk, double m,
                                double h, double rho, double C_free) {
    if (same) {
        if (p.f[ot] == 0.0) return (h * h) * p.f[oL] + C_free * std::max(p.f[olam], 0.0);
        else return (h * h) * p.f[oL];
    }
    return 2.0 * (h * h) * m / rho * rDk(h, r);
}

}  // namespace

// ============================================================== C API (ctypes)
extern "C" {

void* so_create(const double lo[3], const double hi[3], double h) {
    if (!(h > 0.0)) return nullptr;  // structs.jl:59
    OSys* s = new OSys();
    s->h = h;
    for (int a = 0; a < 3; a++) {
        s->lo[a] = lo[a];
        s->hi[a] = hi[a];
    }
    init_keys(*s);
    return s;
}
void so_destroy(void* hnd) { delete (OSys*)hnd; }
int so_num_slots() { return NSLOT; }
int so_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void so_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void so_key_params(void* hnd, int64_t phase[3], int64_t lim[3], int64_t* key_max, int32_t* ndiff, int64_t diff[27]) {
    OSys& s = *(OSys*)hnd;
    for (int a = 0; a < 3; a++) {
        phase[a] = s.key_phase[a];
        lim[a] = s.key_lim[a];
    }
    *key_max = s.key_max;
    *ndiff = (int32_t)s.key_diff.size();
    for (size_t i = 0; i < s.key_diff.size(); i++) diff[i] = s.key_diff[i];
}

void so_resize(void* hnd, int64_t n) {
    OSys& s = *(OSys*)hnd;
    Particle z;
    std::memset(&z, 0, sizeof z);
    s.particles.resize(n, z);
    s.have_cells = false;
}
int64_t so_num_particles(void* hnd) { return (int64_t)((OSys*)hnd)->particles.size(); }
int64_t so_num_removed(void* hnd) { return ((OSys*)hnd)->n_removed; }

// host AoS [n][ncomp] <-> record offset `slot`
void so_set_field(void* hnd, int slot, int ncomp, const double* host) {
    OSys& s = *(OSys*)hnd;
    const int64_t N = (int64_t)s.particles.size();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++)
        for (int c = 0; c < ncomp; c++) s.particles[i].f[slot + c] = host[i * ncomp + c];
}
void so_get_field(void* hnd, int slot, int ncomp, double* host) {
    OSys& s = *(OSys*)hnd;
    const int64_t N = (int64_t)s.particles.size();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++)
        for (int c = 0; c < ncomp; c++) host[i * ncomp + c] = s.particles[i].f[slot + c];
}

void so_create_cell_list(void* hnd) { create_cell_list(*(OSys*)hnd); }

int so_apply(void* hnd, int op, const int32_t* fields, int nf, const double* params, int np, int flags) {
    OSys& s = *(OSys*)hnd;
    return apply_op(s, op, fields, nf, params, np, flags);
}

// keys[i] = find_key(particles[i].x)
void so_get_cell_keys(void* hnd, int64_t* keys) {
    OSys& s = *(OSys*)hnd;
    for (size_t i = 0; i < s.particles.size(); i++) keys[i] = find_key(s, s.particles[i].f);
}
void so_get_cell_list(void* hnd, int64_t* offsets, int64_t* members) {
    OSys& s = *(OSys*)hnd;
    std::copy(s.cell_start.begin(), s.cell_start.end(), offsets);
    std::copy(s.entries.begin(), s.entries.end(), members);
}
// returns 1 if the literal insertion path (core.jl:13-41) yields exactly the CSR cell list
int so_check_cell_list_literal(void* hnd) {
    OSys& s = *(OSys*)hnd;
    std::vector<std::vector<int64_t>> cells;
    create_cell_list_literal(s, cells);
    for (int64_t k = 1; k <= s.key_max; k++) {
        const auto& e = cells[k - 1];
        int64_t n = s.cell_start[k] - s.cell_start[k - 1];
        size_t cnt = 0;
        while (cnt < e.size() && e[cnt] != 0) cnt++;
        if ((int64_t)cnt != n) return 0;
        for (int64_t t = 0; t < n; t++)
            if (e[t] != s.entries[s.cell_start[k - 1] + t]) return 0;
    }
    return 1;
}

// neighbour lists in the visiting order of _apply_binary! (core.jl:94-112)
int64_t so_get_neighbour_lists(void* hnd, int64_t* offsets, int64_t* ids, int64_t cap) {
    OSys& s = *(OSys*)hnd;
    const int64_t N = (int64_t)s.particles.size();
    int64_t total = 0;
    for (int64_t i = 1; i <= N; i++) {
        offsets[i - 1] = total;
        Particle& p = s.particles[i - 1];
        int64_t key = find_key(s, p.f);
        for (int64_t dkey : s.key_diff) {
            int64_t nk = key + dkey;
            if (1 <= nk && nk <= s.key_max)
                for (int64_t e = s.cell_start[nk - 1]; e < s.cell_start[nk]; e++) {
                    int64_t j = s.entries[e];
                    double d[3];
                    double r = dist3(p.f, s.particles[j - 1].f, d);
                    if (r > s.h || j == i) continue;
                    if (ids && total < cap) ids[total] = j;
                    total++;
                }
        }
    }
    offsets[N] = total;
    return total;
}

// sum(sys, func, x), core.jl:240-260 for the registered point sums (cavity_flow.jl:162-180)
int so_sum_at_points(void* hnd, int sum_op, const int32_t* F, int nf, const double* P, int np, const double* xyz,
                     int64_t m_pts, double* out) {
    OSys& s = *(OSys*)hnd;
    if (!((sum_op == SP_SUM_MASS_W && nf == 2 && np == 4) || (sum_op == SP_SUM_MASS_F_W && nf == 3 && np == 5)))
        return SP_ERR_INVALID;
    kfn w = pick_w((int)P[0]);
    if (!w) return SP_ERR_INVALID;
    const double m = P[1], h = P[2], tsel = P[3];
    const int ot = F[1];
    const int of = sum_op == SP_SUM_MASS_F_W ? F[2] + (int)P[4] : 0;
    for (int64_t k = 0; k < m_pts; k++) {
        const double* x = xyz + 3 * k;
        double acc = 0.0;
        int64_t key = find_key(s, x);
        for (int64_t dkey : s.key_diff) {
            int64_t nk = key + dkey;
            if (1 <= nk && nk <= s.key_max)
                for (int64_t e = s.cell_start[nk - 1]; e < s.cell_start[nk]; e++) {
                    const Particle& q = s.particles[s.entries[e] - 1];
                    double d[3];
                    double r = dist3(x, q.f, d);  // norm(x - q.x)
                    if (r > s.h) continue;
                    double sel = (q.f[ot] == tsel) ? 1.0 : 0.0;  // Float64(p.type==FLUID)
                    if (sum_op == SP_SUM_MASS_W) acc += sel * m * w(h, r);
                    else acc += sel * m * q.f[of] * w(h, r);
                }
        }
        out[k] = acc;
    }
    return SP_OK;
}

// serial diagnostic loops of the examples
int so_reduce(void* hnd, int red, const int32_t* F, int nf, const double* P, int np, double* out) {
    OSys& s = *(OSys*)hnd;
    switch (red) {
        case SP_RED_ENERGY_WCSPH: {  // collapse_dry.jl:166-171
            if (nf != 3 || np != 6) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1], orho = F[2];
            const double m = P[0], c = P[1], rho0 = P[2], g[3] = {P[3], P[4], P[5]};
            double E = 0.0;
            for (const Particle& p : s.particles) {
                double kinetic = 0.5 * m * dot3(p.f + ov, p.f + ov);
                double potential = -m * dot3(g, p.f + ox);
                double internal = m * (c * c) * (std::log(std::fabs(p.f[orho] / rho0)) + rho0 / p.f[orho] - 1.0);
                E += kinetic + potential + internal;
            }
            out[0] = E;
            return SP_OK;
        }
        case SP_RED_FRONT: {  // collapse_dry.jl:173-187
            if (nf != 2 || np != 4) return SP_ERR_INVALID;
            const int ox = F[0], ot = F[1];
            const double width = P[0], height = P[1], h = P[2], xmax = P[3];
            double X = 0.0, H = 0.0;
            for (const Particle& p : s.particles) {
                if (p.f[ot] == 0.0) X = std::max(X, p.f[ox] / width);
                if (p.f[ot] == 0.0 && xmax > p.f[ox] && p.f[ox] > h) H = std::max(H, p.f[ox + 1] / height);
            }
            out[0] = X;
            out[1] = H;
            return SP_OK;
        }
        case SP_RED_ENERGY_COLLISION: {  // test_collision_2d.jl:96-100
            if (nf != 3 || np != 3) return SP_ERR_INVALID;
            const int ov = F[0], orho = F[1], orho0 = F[2];
            const double m = P[0], c = P[1], rho0 = P[2];
            double E = 0.0;
            for (const Particle& p : s.particles) {
                double kinetic = 0.5 * m * dot3(p.f + ov, p.f + ov);
                double dr = p.f[orho] - p.f[orho0];
                double internal = 0.5 * m * (c * c) * (dr * dr) / (rho0 * rho0);
                E += kinetic + internal;
            }
            out[0] = E;
            return SP_OK;
        }
        case SP_RED_SUM: {
            if (nf != 2) return SP_ERR_INVALID;
            for (int c = 0; c < F[1]; c++) {
                double a = 0.0;
                for (const Particle& p : s.particles) a += p.f[F[0] + c];
                out[c] = a;
            }
            return SP_OK;
        }
        case SP_RED_ENERGY_ISPH: {  // collapse_dry_implicit.jl:173-177
            if (nf != 2 || np != 4) return SP_ERR_INVALID;
            const int ox = F[0], ov = F[1];
            const double m = P[0], g[3] = {P[1], P[2], P[3]};
            double E = 0.0;
            for (const Particle& p : s.particles)
                E += 0.5 * m * dot3(p.f + ov, p.f + ov) + -m * dot3(g, p.f + ox);
            out[0] = E;
            return SP_OK;
        }
        case SP_RED_FORCE_ON_TYPE: {  // cylinder.jl:158-159: F = sum(p -> p.m*p.a, obstacle)
            if (nf != 3 || np != 1) return SP_ERR_INVALID;
            const int oa = F[0], om = F[1], ot = F[2];
            double acc[3] = {0.0, 0.0, 0.0};
            for (const Particle& p : s.particles)
                if (p.f[ot] == P[0])
                    for (int c = 0; c < 3; c++) acc[c] += p.f[om] * p.f[oa + c];
            out[0] = acc[0];
            out[1] = acc[1];
            out[2] = acc[2];
            return SP_OK;
        }
        case SP_RED_MAX_SPEED: {  // max norm(v) (algebra.jl:58-60); no counterpart in the reference
            if (nf != 1 || np != 0) return SP_ERR_INVALID;
            double vmax = 0.0;
            for (const Particle& p : s.particles) vmax = std::max(vmax, std::sqrt(dot3(p.f + F[0], p.f + F[0])));
            out[0] = vmax;
            return SP_OK;
        }
        case SP_RED_ENERGY_ROD: {  // rod.jl:190-199 (particle_energy summed as in :213)
            if (nf != 2 || np != 3) return SP_ERR_INVALID;
            const int ov = F[0], oA = F[1];
            const double m = P[0], c_s = P[1], c_l = P[2];
            double E = 0.0;
            for (const Particle& p : s.particles) {
                M2 A = m2_load(p.f + oA);
                double d = std::fabs(m2_det(A));
                double lam;
                M2 G0 = m2_dev(m2_mul(m2_trans(A), A), &lam);
                double g33 = 1.0 - lam;
                // LinearAlgebra.norm(G0, 2) of an SMatrix: sqrt of the sum of squares in linear index order
                double n2 = std::sqrt(G0.a11 * G0.a11 + G0.a21 * G0.a21 + G0.a12 * G0.a12 + G0.a22 * G0.a22 + g33 * g33);
                double E_kinet = 0.5 * m * dot3(p.f + ov, p.f + ov);
                double E_shear = 0.25 * m * (c_s * c_s) * (n2 * n2);
                double E_press = m * (c_l * c_l) * (d - 1.0 - std::log(d));
                E += E_kinet + E_shear + E_press;
            }
            out[0] = E;
            return SP_OK;
        }
    }
    return SP_ERR_INVALID;
}

// assemble_matrix(sys, projection_matrix), core.jl:196-225: COO triplets in visiting order, 1-based,
// INCLUDING the diagonal (no p == q skip).  Returns nnz; fills arrays when non-null and cap suffices.
// fields {x, L, lambda, type}; params {kernel, m, h, rho, C_free}
int64_t so_assemble_matrix(void* hnd, const int32_t* F, int nf, const double* P, int np, int64_t* I, int64_t* J,
                           double* V, int64_t cap) {
    OSys& s = *(OSys*)hnd;
    if (nf < 4 || np != 5) return -1;
    const int oL = F[1], olam = F[2], ot = F[3];
    kfn rDk = pick_rD((int)P[0]);
    if (!rDk) return -1;
    const double m = P[1], h = P[2], rho = P[3], C_free = P[4];
    const int64_t N = (int64_t)s.particles.size();
    int64_t nnz = 0;
    for (int64_t i = 1; i <= N; i++) {
        const Particle& p = s.particles[i - 1];
        int64_t key = find_key(s, p.f);
        for (int64_t dkey : s.key_diff) {
            int64_t nk = key + dkey;
            if (1 <= nk && nk <= s.key_max)
                for (int64_t e = s.cell_start[nk - 1]; e < s.cell_start[nk]; e++) {
                    int64_t l = s.entries[e];
                    double d[3];
                    double r = dist3(p.f, s.particles[l - 1].f, d);
                    if (r > s.h) continue;  // :214
                    if (I && nnz < cap) {
                        I[nnz] = i;
                        J[nnz] = l;
                        V[nnz] = projection_matrix(p, l == i, r, oL, olam, ot, rDk, m, h, rho, C_free);
                    }
                    nnz++;
                }
        }
    }
    return nnz;
}

// y = A x for COO triplets (duplicates add, like sparse(I,J,V), core.jl:224)
void so_coo_matvec(int64_t nnz, const int64_t* I, const int64_t* J, const double* V, const double* x, double* y,
                   int64_t n) {
    for (int64_t i = 0; i < n; i++) y[i] = 0.0;
    for (int64_t k = 0; k < nnz; k++) y[I[k] - 1] += V[k] * x[J[k] - 1];
}

// cg(A, b): IterativeSolvers.jl (third-party, NOT under /root/reference, version unpinned — the package's
// Project.toml does not declare it; call site examples/collapse_dry_implicit.jl:40,227).  Published
// algorithm of its un-preconditioned CGIterable: x0 = 0, u = 0, r = b, tol = max(reltol*|r0|, abstol);
// loop while |r| > tol and it < maxiter: beta = |r|^2/|r_prev|^2 (|r_prev| = 1 initially);
// u = r + beta u; c = A u; alpha = |r|^2 / (u.c); x += alpha u; r -= alpha c.
int64_t so_cg_coo(int64_t nnz, const int64_t* I, const int64_t* J, const double* V, const double* b, double* x,
                  int64_t n, double reltol, double abstol, int64_t maxiter, double* resid_out) {
    std::vector<double> r(b, b + n), u(n, 0.0), c(n, 0.0);
    for (int64_t i = 0; i < n; i++) x[i] = 0.0;
    auto nrm = [&](const std::vector<double>& v) {
        double a = 0.0;
        for (double t : v) a += t * t;
        return std::sqrt(a);
    };
    double residual = nrm(r), prev = 1.0;
    const double tol = std::max(reltol * residual, abstol);
    if (maxiter <= 0) maxiter = n;
    int64_t it = 0;
    while (it < maxiter && residual > tol) {
        double beta = residual * residual / (prev * prev);
        for (int64_t i = 0; i < n; i++) u[i] = r[i] + beta * u[i];
        so_coo_matvec(nnz, I, J, V, u.data(), c.data(), n);
        double uc = 0.0;
        for (int64_t i = 0; i < n; i++) uc += u[i] * c[i];
        double alpha = residual * residual / uc;
        for (int64_t i = 0; i < n; i++) {
            x[i] += alpha * u[i];
            r[i] -= alpha * c[i];
        }
        prev = residual;
        residual = nrm(r);
        it++;
    }
    if (resid_out) *resid_out = residual;
    return it;
}

// add_new_particles!, cylinder.jl:145-156 (literal, serial): slot/value pairs describe what the script's constructor
// Particle(x, INFLOW) sets besides x and type (every other field is zero).
int64_t so_respawn(void* hnd, int type_slot, double from_type, double to_type, double x1_min, double shift,
                   const int32_t* fill_slots, const double* fill_values, int n_fill) {
    OSys& s = *(OSys*)hnd;
    std::vector<Particle> fresh;
    for (Particle& p : s.particles) {
        if (p.f[type_slot] == from_type && p.f[0] >= x1_min) {
            p.f[type_slot] = to_type;
            Particle q;
            for (int k = 0; k < NSLOT; k++) q.f[k] = 0.0;
            q.f[0] = p.f[0] - shift * 1.0;  // p.x - bc_width*VECX
            q.f[1] = p.f[1] - shift * 0.0;
            q.f[2] = p.f[2] - shift * 0.0;
            for (int k = 0; k < n_fill; k++) q.f[fill_slots[k]] = fill_values[k];
            q.f[type_slot] = from_type;
            fresh.push_back(q);
        }
    }
    s.particles.insert(s.particles.end(), fresh.begin(), fresh.end());
    s.have_cells = false;
    return (int64_t)fresh.size();
}

void so_kernel_eval(int kernel, int kfun, double h, const double* r, double* out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = kernel_eval(kernel, kfun, h, r[i]);
}

// Step programs for the CPU baseline: the time loops of the examples without I/O.
// fields {x, v, Dv, rho, Drho, P, type}; params {kernel, m, h, two_nu, dt, c2, rho0, mu, gx, gy, gz}
// returns wall seconds of the loop.
double so_run_program(void* hnd, int program, const int32_t* F, int nf, const double* P, int np, int64_t nsteps) {
    OSys& s = *(OSys*)hnd;
    if (nf != 7 || np != 11) return -1.0;
    const int32_t x = F[0], v = F[1], Dv = F[2], rho = F[3], Drho = F[4], Pr = F[5], ty = F[6];
    const double kernel = P[0], m = P[1], h = P[2], two_nu = P[3], dt = P[4], c2 = P[5], rho0 = P[6], mu = P[7];
    const int32_t f_bom[4] = {x, v, rho, Drho}, f_fp[3] = {rho, Drho, Pr}, f_if[6] = {x, v, Pr, rho, Dv, ty},
                  f_mv[4] = {x, v, Dv, ty}, f_ac[3] = {v, Dv, ty};
    const double p_bom[4] = {kernel, m, h, two_nu}, p_fp[4] = {dt, c2, rho0, 0.0}, p_if[5] = {kernel, m, h, mu, rho0},
                 p_ac[4] = {0.5 * dt, P[8], P[9], P[10]};
    auto t0 = std::chrono::steady_clock::now();
    for (int64_t k = 0; k < nsteps; k++) {
        if (program == SP_PROGRAM_WCSPH_3D) {  // collapse3d.jl:136-150
            const double p_mv[1] = {dt};
            apply_op(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0);
            create_cell_list(s);
            apply_op(s, SP_OP_BALANCE_OF_MASS, f_bom, 4, p_bom, 4, 0);
            apply_op(s, SP_OP_FIND_PRESSURE, f_fp, 3, p_fp, 4, 0);
            apply_op(s, SP_OP_INTERNAL_FORCE, f_if, 6, p_if, 5, 0);
            apply_op(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0);
            apply_op(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0);
        } else if (program == SP_PROGRAM_WCSPH_2D) {  // collapse_dry.jl:203-211
            const double p_mv[1] = {0.5 * dt};
            apply_op(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0);
            apply_op(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0);
            create_cell_list(s);
            apply_op(s, SP_OP_BALANCE_OF_MASS, f_bom, 4, p_bom, 4, 0);
            apply_op(s, SP_OP_FIND_PRESSURE, f_fp, 3, p_fp, 4, 0);
            apply_op(s, SP_OP_MOVE, f_mv, 4, p_mv, 1, 0);
            create_cell_list(s);
            apply_op(s, SP_OP_INTERNAL_FORCE, f_if, 6, p_if, 5, 0);
            apply_op(s, SP_OP_ACCELERATE, f_ac, 3, p_ac, 4, 0);
        } else
            return -1.0;
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
