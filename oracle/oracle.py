"""ctypes wrapper of the CPU oracle (oracle/sp_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

``OracleSystem`` exposes the same methods as ``smoothedparticles_jl_b200.ParticleSystem`` so a parity test
runs one "program" (a list of apply / create_cell_list calls) on both and compares the fields.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Mapping, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libsp_oracle.so")
LIB_PATH_WIDE = os.path.join(_HERE, "_build", "libsp_oracle_wide.so")  # 48-slot particle record (rod.jl)
LIB_PATH_FAST = os.path.join(_HERE, "_build", "libsp_oracle_fast.so")  # -ffast-math build: a yardstick, not an oracle
_lib = None
_lib_wide = None

_p, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pi32, _pi64, _pf64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sp_oracle.cpp")
    hdr = os.path.join(os.path.dirname(_HERE), "include", "sp_b200.h")
    if not force and all(os.path.exists(q) and os.path.getmtime(q) >= os.path.getmtime(src)
                         and os.path.getmtime(q) >= os.path.getmtime(hdr) for q in (LIB_PATH, LIB_PATH_WIDE, LIB_PATH_FAST)):
        return LIB_PATH
    r = subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def load(wide: bool = False):
    global _lib, _lib_wide
    if wide and _lib_wide is not None:
        return _lib_wide
    if not wide and _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH_WIDE if wide else LIB_PATH)
    sig = {
        "so_create": (_p, [_pf64, _pf64, _f64]),
        "so_destroy": (None, [_p]),
        "so_num_slots": (C.c_int, []),
        "so_max_threads": (C.c_int, []),
        "so_set_threads": (None, [C.c_int]),
        "so_key_params": (None, [_p, _pi64, _pi64, _pi64, _pi32, _pi64]),
        "so_resize": (None, [_p, _i64]),
        "so_num_particles": (_i64, [_p]),
        "so_num_removed": (_i64, [_p]),
        "so_set_field": (None, [_p, C.c_int, C.c_int, _pf64]),
        "so_get_field": (None, [_p, C.c_int, C.c_int, _pf64]),
        "so_create_cell_list": (None, [_p]),
        "so_apply": (C.c_int, [_p, C.c_int, _pi32, C.c_int, _pf64, C.c_int, C.c_int]),
        "so_get_cell_keys": (None, [_p, _pi64]),
        "so_get_cell_list": (None, [_p, _pi64, _pi64]),
        "so_check_cell_list_literal": (C.c_int, [_p]),
        "so_get_neighbour_lists": (_i64, [_p, _pi64, _pi64, _i64]),
        "so_sum_at_points": (C.c_int, [_p, C.c_int, _pi32, C.c_int, _pf64, C.c_int, _pf64, _i64, _pf64]),
        "so_reduce": (C.c_int, [_p, C.c_int, _pi32, C.c_int, _pf64, C.c_int, _pf64]),
        "so_assemble_matrix": (_i64, [_p, _pi32, C.c_int, _pf64, C.c_int, _pi64, _pi64, _pf64, _i64]),
        "so_coo_matvec": (None, [_i64, _pi64, _pi64, _pf64, _pf64, _pf64, _i64]),
        "so_cg_coo": (_i64, [_i64, _pi64, _pi64, _pf64, _pf64, _pf64, _i64, _f64, _f64, _i64, _pf64]),
        "so_kernel_eval": (None, [C.c_int, C.c_int, _f64, _pf64, _pf64, _i64]),
        "so_respawn": (_i64, [_p, C.c_int, _f64, _f64, _f64, _f64, _pi32, _pf64, C.c_int]),
        "so_run_program": (_f64, [_p, C.c_int, _pi32, C.c_int, _pf64, C.c_int, _i64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if wide:
        _lib_wide = lib
    else:
        _lib = lib
    return lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pf(a):
    return a.ctypes.data_as(_pf64)


def _pi(a):
    return a.ctypes.data_as(_pi64)


def _p32(a):
    return a.ctypes.data_as(_pi32)


FILL_OP = 10  # SP_OP_FILL takes {slot, ncomp} on the oracle side


class OracleSystem:
    def __init__(self, particle_fields: Mapping[str, int], domain, h: float, threads: int | None = None):
        self._lib = load(wide=3 + sum(int(nc) for nm, nc in particle_fields.items() if nm != "x") > 20)
        box = domain.boundarybox() if hasattr(domain, "boundarybox") else domain
        lo = (C.c_double * 3)(*[float(v) for v in box.lo])
        hi = (C.c_double * 3)(*[float(v) for v in box.hi])
        self._h = self._lib.so_create(lo, hi, float(h))
        if not self._h:
            raise ValueError("invalid ParticleSystem declaration! (h must be a positive float)")
        if threads:
            self._lib.so_set_threads(int(threads))
        self.h = float(h)
        self.fields: Dict[str, int] = {"x": 3}
        self._slot: Dict[str, int] = {"x": 0}
        self._next = 3
        for name, nc in particle_fields.items():
            self.add_field(name, nc)
        phase = (C.c_int64 * 3)()
        lim = (C.c_int64 * 3)()
        kmax = C.c_int64()
        nd = C.c_int32()
        diff = (C.c_int64 * 27)()
        self._lib.so_key_params(self._h, phase, lim, C.byref(kmax), C.byref(nd), diff)
        self.key_phase, self.key_lim, self.key_max = tuple(phase), tuple(lim), kmax.value
        self.key_diff = list(diff[: nd.value])

    def close(self):
        if getattr(self, "_h", None):
            self._lib.so_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_field(self, name, ncomp=1):
        if name in self._slot:
            return self._slot[name]
        if self._next + ncomp > self._lib.so_num_slots():
            raise ValueError("oracle particle record is full")
        self.fields[name] = ncomp
        self._slot[name] = self._next
        self._next += ncomp
        return self._slot[name]

    def __len__(self):
        return self._lib.so_num_particles(self._h)

    @property
    def n(self):
        return len(self)

    @property
    def n_removed(self):
        return self._lib.so_num_removed(self._h)

    def resize(self, n):
        self._lib.so_resize(self._h, int(n))

    def add_particles(self, **arrays):
        x = _f(arrays["x"])
        n_old, n_new = len(self), x.shape[0]
        if n_new == 0:
            return
        cur = {nm: self.get(nm) for nm in self.fields} if n_old else {}
        self.resize(n_old + n_new)
        for name, nc in self.fields.items():
            old = cur.get(name, np.zeros((0, nc) if nc > 1 else (0,)))
            if name in arrays:
                add = np.broadcast_to(_f(arrays[name]), (n_new,) if nc == 1 else (n_new, nc))
            else:
                add = np.zeros((n_new,) if nc == 1 else (n_new, nc))
            self.set(name, np.concatenate([old, add], axis=0))

    def respawn(self, type_field, from_type, to_type, x1_min, shift, **constants):
        """add_new_particles!, examples/cylinder.jl:145-156 (literal serial loop in the oracle)."""
        names = list(constants)
        slots = np.asarray([self._slot[n] for n in names] or [0], dtype=np.int32)
        vals = _f([constants[n] for n in names] or [0.0])
        return int(self._lib.so_respawn(self._h, self._slot[type_field], float(from_type), float(to_type),
                                        float(x1_min), float(shift), _p32(slots), _pf(vals), len(names)))

    def set(self, name, values):
        a = _f(values)
        nc = self.fields[name]
        if a.size != len(self) * nc:
            raise ValueError("size mismatch")
        self._lib.so_set_field(self._h, self._slot[name], nc, _pf(a))

    def get(self, name):
        nc = self.fields[name]
        n = len(self)
        out = np.empty((n, nc) if nc > 1 else (n,))
        if n:
            self._lib.so_get_field(self._h, self._slot[name], nc, _pf(out))
        return out

    def create_cell_list(self):
        self._lib.so_create_cell_list(self._h)

    def _bind(self, names):
        return np.asarray([self._slot[nm] for nm in names], dtype=np.int32)

    def apply(self, op, self_=False, strict_order=False):
        if op.op == FILL_OP:
            F = np.asarray([self._slot[op.fields[0]], self.fields[op.fields[0]]], dtype=np.int32)
        else:
            F = self._bind(op.fields)
        P = _f(op.params)
        rc = self._lib.so_apply(self._h, op.op, _p32(F), len(F), _pf(P), len(P), 1 if self_ else 0)
        if rc != 0:
            raise RuntimeError(f"oracle: operator {op.op} rejected (status {rc})")

    def sum_at_points(self, sum_op, fields, params, points):
        pts = _f(points).reshape(-1, 3)
        out = np.empty(len(pts))
        F = self._bind(fields)
        if len(F) == 3:  # the third binding is a base slot; the component is in params
            pass
        P = _f(params)
        rc = self._lib.so_sum_at_points(self._h, sum_op, _p32(F), len(F), _pf(P), len(P), _pf(pts), len(pts), _pf(out))
        if rc != 0:
            raise RuntimeError("oracle: point sum rejected")
        return out

    def reduce(self, red, fields, params=(), nout=1):
        if red == 4:  # SP_RED_SUM takes {slot, ncomp}
            F = np.asarray([self._slot[fields[0]], self.fields[fields[0]]], dtype=np.int32)
        else:
            F = self._bind(fields)
        P = _f(params) if len(params) else np.zeros(1)
        out = np.zeros(3)
        rc = self._lib.so_reduce(self._h, red, _p32(F), len(F), _pf(P), len(params), _pf(out))
        if rc != 0:
            raise RuntimeError("oracle: reduction rejected")
        return out[:nout].copy()

    def assemble_vector(self, op):
        self.apply(op)
        return self.get(op.fields[-1])

    def assemble_matrix(self, A):
        """COO triplets (I, J, V), 1-based, in the reference's visiting order (src/core.jl:196-225)."""
        F = self._bind(A.fields)
        P = _f(A.params)
        nnz = self._lib.so_assemble_matrix(self._h, _p32(F), len(F), _pf(P), len(P), None, None, None, 0)
        if nnz < 0:
            raise RuntimeError("oracle: assemble_matrix rejected")
        I = np.empty(nnz, dtype=np.int64)
        J = np.empty(nnz, dtype=np.int64)
        V = np.empty(nnz)
        self._lib.so_assemble_matrix(self._h, _p32(F), len(F), _pf(P), len(P), _pi(I), _pi(J), _pf(V), nnz)
        return I, J, V

    def cg(self, I, J, V, b, reltol=None, abstol=0.0, maxiter=0):
        if reltol is None:
            reltol = float(np.sqrt(np.finfo(np.float64).eps))
        b = _f(b)
        x = np.empty_like(b)
        resid = C.c_double()
        it = self._lib.so_cg_coo(len(V), _pi(I), _pi(J), _pf(V), _pf(b), _pf(x), len(b), reltol, abstol, maxiter,
                                 C.byref(resid))
        return x, it, resid.value

    def poisson_cg(self, A, b, P_out, reltol=None, abstol=0.0, maxiter=0):
        """The reference's own path for ``A = assemble_matrix(sys, projection_matrix); P .= cg(A, b)``
        (collapse_dry_implicit.jl:223-227): serial assembly of the COO triplets, then CG on them; same signature and
        return value (iterations, residual norm) as ParticleSystem.poisson_cg so that Case.step runs on both."""
        I, J, V = self.assemble_matrix(A)
        x, it, resid = self.cg(I, J, V, self.get(b), reltol=reltol, abstol=abstol, maxiter=maxiter)
        self.set(P_out, x)
        return it, resid

    def coo_matvec(self, I, J, V, x):
        x = _f(x)
        y = np.empty_like(x)
        self._lib.so_coo_matvec(len(V), _pi(I), _pi(J), _pf(V), _pf(x), _pf(y), len(x))
        return y

    def run_program(self, program, fields, params, nsteps) -> float:
        F = self._bind(fields)
        P = _f(params)
        t = self._lib.so_run_program(self._h, program, _p32(F), len(F), _pf(P), len(P), int(nsteps))
        if t < 0:
            raise RuntimeError("oracle: program rejected")
        return t

    # parity views
    def cell_keys(self):
        out = np.empty(len(self), dtype=np.int64)
        self._lib.so_get_cell_keys(self._h, _pi(out))
        return out

    def cell_list(self):
        offsets = np.empty(self.key_max + 1, dtype=np.int64)
        members = np.empty(max(len(self), 1), dtype=np.int64)
        self._lib.so_get_cell_list(self._h, _pi(offsets), _pi(members))
        return offsets, members[: len(self)]

    def check_cell_list_literal(self) -> bool:
        return bool(self._lib.so_check_cell_list_literal(self._h))

    def neighbour_lists(self):
        n = len(self)
        offsets = np.zeros(n + 1, dtype=np.int64)
        total = self._lib.so_get_neighbour_lists(self._h, _pi(offsets), None, 0)
        ids = np.empty(max(total, 1), dtype=np.int64)
        self._lib.so_get_neighbour_lists(self._h, _pi(offsets), _pi(ids), total)
        return offsets, ids[:total]


def kernel_eval_fastmath(kernel_id: int, kfun: int, h: float, r):
    """The kernel functions as a compiler that may reassociate / contract / use reciprocals evaluates them (gcc
    -ffast-math -mfma): how far the reference's own @fastmath kernels (src/kernels.jl) may sit from the strictly
    rounded restatement.  Used only to size the tolerance, never as a reference value."""
    build()
    lib = C.CDLL(LIB_PATH_FAST)
    lib.so_kernel_eval.restype = None
    lib.so_kernel_eval.argtypes = [C.c_int, C.c_int, _f64, _pf64, _pf64, _i64]
    r = _f(r)
    out = np.empty_like(r)
    lib.so_kernel_eval(int(kernel_id), int(kfun), float(h), _pf(r), _pf(out), r.size)
    return out


def kernel_eval(kernel_id: int, kfun: int, h: float, r):
    lib = load()
    r = _f(r)
    out = np.empty_like(r)
    lib.so_kernel_eval(int(kernel_id), int(kfun), float(h), _pf(r), _pf(out), r.size)
    return out


def max_threads() -> int:
    return load().so_max_threads()
