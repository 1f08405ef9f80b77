"""Runs the DEVICE operator bodies on the host: a second look at the operators without a GPU (every operator also has a B200 parity test).

The operator structs of csrc/sp_ops.cuh (and the kernel families of csrc/sp_kernels.cuh) are plain C++ apart from the
CUDA qualifiers and a handful of rounding intrinsics, so they are compiled for the host with those mapped to their
IEEE meaning (`__dmul_rn(a, b)` -> `a*b` under -ffp-contract=off, ...).  The field/parameter binding is not rewritten by
hand either: the `case SP_OP_...` blocks of `sp_apply_impl` (csrc/sp_sweep.cu) are extracted from the source TEXT and
transliterated mechanically (`sc(s, F[k])` -> the k-th bound field, `dispatch_kernel<Op>` -> a sequential sweep over
explicit neighbour lists, `launch_unary<U>` -> a loop), so a wrong plane, a swapped parameter or a sign error in what
was written for the GPU shows up here, against the oracle, without a GPU.  What this cannot see: the CUDA kernels
around the operators (list build, replay, cell list) — those are exercised by the GPU parity tests of the operators
that have already run on the B200.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "smoothedparticles.jl_b200", "csrc")
BUILD = os.path.join(ROOT, "oracle", "_build", "host_ops")

PREAMBLE = r'''
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include "sp_b200.h"
using std::isfinite; using std::fabs; using std::sqrt; using std::log; using std::fmax;
#define __CUDACC__ 1
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __dadd_rn(a, b) ((a) + (b))
#define __dsub_rn(a, b) ((a) - (b))
#define __dmul_rn(a, b) ((a) * (b))
#define __fma_rn(a, b, c) std::fma((a), (b), (c))
#define __double2ll_rn(x) ((long long)std::nearbyint(x))
#define __ll2double_rn(x) ((double)(x))
'''

HARNESS = r'''
struct Ctx {
    double** fld;        // bound fields in the operator's documented order (component-major planes, stride n)
    long long n;
    const long long *off, *ids;   // neighbour lists in the reference's visiting order, 1-based ids
    int self;
    double* sc(int k) const { return fld[k]; }
    RV3 rv3(int k) const { return RV3{fld[k], fld[k] + n, fld[k] + 2 * n}; }
    WV3 wv3(int k) const { return WV3{fld[k], fld[k] + n, fld[k] + 2 * n}; }
    void set_v3(int k, const double** qp) const { qp[0] = fld[k]; qp[1] = fld[k] + n; qp[2] = fld[k] + 2 * n; }
    double* pressure_over_rho2(int kP, int krho) const {   // pressure_over_rho2() of sp_sweep.cu: pr = P/rho^2 per particle
        double* pr = (double*)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));   // leaked on purpose: test process
        UPressureOverRho2::Params Pp{fld[kP], fld[krho], pr};
        for (long long i = 0; i < n; i++) UPressureOverRho2::apply(Pp, (int)i);
        return pr;
    }
};
template <int NQ>
struct QHost {
    const double* const* qp;
    long long j;
    double operator()(int k) const { return qp[k][j]; }
};
template <class Op>
static int sweep(const Ctx& c, typename Op::Params& P) {   // k_sweep<Op, true>: sequential, visiting order, exact sqrt
    const double *X = c.fld[0], *Y = c.fld[0] + c.n, *Z = c.fld[0] + 2 * c.n;
    for (long long i = 0; i < c.n; i++) {
        if (!Op::active(P, (int)i)) continue;
        typename Op::PS p;
        typename Op::Acc acc;
        Op::load(P, (int)i, X[i], Y[i], Z[i], p, acc);
        for (long long e = c.off[i]; e < c.off[i + 1]; e++) {
            const long long j = c.ids[e] - 1;
            const double dx = X[i] - X[j], dy = Y[i] - Y[j], dz = Z[i] - Z[j];
            QHost<Op::NQ> q{P.qp, j};
            Op::pair(P, p, q, dx, dy, dz, std::sqrt((dx * dx + dy * dy) + dz * dz), acc);
        }
        if (c.self) Op::self(P, p, acc);
        Op::store(P, (int)i, p, acc);
    }
    return 0;
}
template <template <class> class OpT, class Mk>
static int host_pair(const Ctx& c, int kernel, double h, Mk&& mk) {
    SpKC kc;
    if (!sp_make_kc(kernel, h, &kc)) return 1;
    switch (kernel) {
        case SP_KERNEL_WENDLAND1: case SP_KERNEL_WENDLAND2: case SP_KERNEL_WENDLAND3: {
            typename OpT<KWendland>::Params P; mk(P); P.kc = kc; return sweep<OpT<KWendland>>(c, P);
        }
        case SP_KERNEL_SPLINE23: { typename OpT<KSpline23>::Params P; mk(P); P.kc = kc; return sweep<OpT<KSpline23>>(c, P); }
        default: { typename OpT<KSpline24>::Params P; mk(P); P.kc = kc; return sweep<OpT<KSpline24>>(c, P); }
    }
}
template <class U>
static int host_unary(const Ctx& c, const typename U::Params& P) {
    for (long long i = 0; i < c.n; i++) U::apply(P, (int)i);
    return 0;
}
extern "C" int ho_apply(int op, double** fld, int nf, const double* Pm, int np, long long n, const long long* off,
                        const long long* ids, int self) {
    Ctx ctx{fld, n, off, ids, self};
    (void)nf; (void)np;
    switch (op) {
@CASES@
    }
    return 2;
}
'''


def _cases(ops):
    src = open(os.path.join(CSRC, "sp_sweep.cu")).read()
    body = src[src.index("int sp_apply_impl("):]
    out = []
    for name in ops:
        m = re.search(r"        case " + name + r": \{\n(.*?)\n        \}\n", body, flags=re.S)
        assert m, name
        txt = m.group(1)
        keep = []
        for line in txt.split("\n"):
            if re.search(r"\bNEED(_CELLS)?\(|sp_wrote\(|sp_zeroed\(|sp_check_fields|if \(rc2\)|np != 0|const int ncs\[\]", line):
                continue
            keep.append(line)
        txt = "\n".join(keep)
        txt = re.sub(r"dispatch_kernel<(\w+)>\(s, ([^,]+), ([^,]+), flags[^,]*, ", r"host_pair<\1>(ctx, \2, \3, ", txt)
        txt = re.sub(r"launch_unary<(\w+)>\(s, P\)", r"host_unary<\1>(ctx, P)", txt)
        txt = re.sub(r"\bsc\(s, F\[(\d+)\]\)", r"ctx.sc(\1)", txt)
        txt = re.sub(r"\brv3\(s, F\[(\d+)\]\)", r"ctx.rv3(\1)", txt)
        txt = re.sub(r"\bwv3\(s, F\[(\d+)\]\)", r"ctx.wv3(\1)", txt)
        txt = re.sub(r"\bset_v3\(s, F\[(\d+)\], ", r"ctx.set_v3(\1, ", txt)
        txt = txt.replace("s->cap", "ctx.n")
        if "s->" in txt or "(s," in txt:
            # cases with scratch-field / cache logic around the sweep: keep the LAST (plain) dispatch statement of the
            # case and emulate the one helper it may depend on, pressure_over_rho2 (the hoisted P/rho^2 plane), with the
            # device's own UPressureOverRho2 body
            at = txt.rfind("return host_pair<")
            assert at >= 0, (name, txt)
            # cases that fill a scratch field of their own in a way other than pressure_over_rho2 (SC_INTERNAL_FORCE: the
            # equation of state; INTERNAL_FORCE_LJ: P/rho^2 or P/rho0^2) are not emulated: the oracle stands in for them
            assert name not in ("SP_OP_SC_INTERNAL_FORCE", "SP_OP_INTERNAL_FORCE_LJ"), name
            hoist = re.search(r"pressure_over_rho2\(s, F\[(\d+)\], F\[(\d+)\], &pr\)", m.group(1))
            txt = (f"            double* pr = ctx.pressure_over_rho2({hoist.group(1)}, {hoist.group(2)});\n" if hoist else "") \
                + "            " + txt[at:]
        assert "s->" not in txt and "(s," not in txt and "flags" not in txt, (name, txt)
        out.append("        case " + name + ": {\n" + txt + "\n        }\n")
    return "".join(out)


_lib = {}


def build(ops):
    key = tuple(ops)
    if key in _lib:
        return _lib[key]
    os.makedirs(BUILD, exist_ok=True)
    for f in ("sp_ops.cuh", "sp_kernels.cuh"):
        txt = open(os.path.join(CSRC, f)).read().replace('#include "sp_internal.cuh"', "")
        open(os.path.join(BUILD, f), "w").write(txt)
    import hashlib
    tag = hashlib.sha1(",".join(ops).encode()).hexdigest()[:10]   # one library per operator set (dlopen caches by path)
    cpp = os.path.join(BUILD, f"harness_{tag}.cpp")
    open(cpp, "w").write(PREAMBLE + '#include "sp_ops.cuh"\n' + HARNESS.replace("@CASES@", _cases(ops)))
    so = os.path.join(BUILD, f"libhost_ops_{tag}.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas",
                        "-I", os.path.join(ROOT, "include"), "-I", BUILD, "-o", so, cpp], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host harness build failed:\n" + r.stderr[-4000:])
    lib = C.CDLL(so)
    lib.ho_apply.restype = C.c_int
    lib.ho_apply.argtypes = [C.c_int, C.POINTER(C.POINTER(C.c_double)), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_longlong,
                             C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]
    _lib[key] = lib
    return lib


class HostFields:
    """Component-major planes of every field of an OracleSystem snapshot (what the device holds, stride n)."""

    def __init__(self, ora, ops, need_lists=True):
        self.lib = build(ops)
        self.n = len(ora)
        self.planes = {}
        for name, nc in ora.fields.items():
            a = ora.get(name).reshape(self.n, nc if nc > 1 else 1)
            self.planes[name] = np.ascontiguousarray(a.T, dtype=np.float64)   # (ncomp, n)
        off, ids = ora.neighbour_lists() if need_lists else (np.zeros(self.n + 1), np.zeros(1))
        self.off = np.ascontiguousarray(off, dtype=np.int64)
        self.ids = np.ascontiguousarray(ids if len(ids) else np.zeros(1), dtype=np.int64)

    def apply(self, op, self_=False):
        ptrs = (C.POINTER(C.c_double) * len(op.fields))(*[self.planes[f].ctypes.data_as(C.POINTER(C.c_double)) for f in op.fields])
        P = np.ascontiguousarray(op.params if len(op.params) else (0.0,), dtype=np.float64)
        rc = self.lib.ho_apply(op.op, ptrs, len(op.fields), P.ctypes.data_as(C.POINTER(C.c_double)), len(op.params), self.n,
                               self.off.ctypes.data_as(C.POINTER(C.c_longlong)), self.ids.ctypes.data_as(C.POINTER(C.c_longlong)),
                               1 if self_ else 0)
        if rc != 0:
            raise RuntimeError(f"host harness: operator {op.op} not built or rejected ({rc})")

    def get(self, name):
        a = self.planes[name].T
        return a[:, 0].copy() if a.shape[1] == 1 else a.copy()


def transliterable_ops():
    """Operator names whose dispatch case is a plain binding (no scratch fields, caches or host-side branches)."""
    import smoothedparticles_jl_b200 as sp
    good = []
    for name in sorted(k for k in sp.K if k.startswith("SP_OP_")):
        try:
            _cases([name])
            good.append(name)
        except AssertionError:
            pass
    return good


def host_backed_system():
    """An OracleSystem whose apply() runs the DEVICE operator body (on the host) for every operator that
    transliterates, and the oracle's own restatement for the few that do not (scratch-field / cache logic in their
    dispatch case); everything else — storage, cell list, neighbour lists, reductions, CG — is the oracle's."""
    import smoothedparticles_jl_b200 as sp
    from oracle.oracle import OracleSystem
    names = transliterable_ops()
    ids = {sp.K[n] for n in names}

    class HostBackedSystem(OracleSystem):
        host_applied = 0

        def apply(self, op, self_=False, strict_order=False):
            if op.op not in ids or len(self) == 0:
                return super().apply(op, self_=self_, strict_order=strict_order)
            host = HostFields(self, names, need_lists=op.binary)
            host.apply(op, self_=self_)
            for f in dict.fromkeys(op.fields):
                self.set(f, host.get(f))
            type(self).host_applied += 1

    return HostBackedSystem
