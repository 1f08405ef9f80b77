"""Host-side mirror of the reference's ``ParticleSystem / create_cell_list! / apply!`` surface
(src/structs.jl:44-125, src/core.jl:51-291) over the C ABI of include/sp_b200.h.

The particle struct of the reference becomes a dict ``{field name: ncomp}``; the particles live in HBM as
struct-of-arrays planes owned by the library.  Per-particle access ``sys.particles[i].x`` becomes bulk
``sys.get("x")`` / ``sys.set("x", array)`` in REFERENCE ORDER (the order the reference's
``sys.particles`` vector would have, including its swap-with-tail renumbering on removal).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Mapping, Sequence

import numpy as np

from . import abi
from .operators import Operator, PoissonOperator

K = abi.K


def _farr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class ParticleSystem:
    """``ParticleSystem(T, domain, h)`` (src/structs.jl:57-91).

    ``particle_fields`` plays the role of the particle type ``T``: field name -> number of components
    (1 scalar, 3 RealVector).  ``x`` (3) always exists.  ``domain`` is any shape with ``boundarybox()``
    (or an object with ``lo``/``hi``); only its bounding box is kept, as in the reference (:63,:87).
    """

    def __init__(self, particle_fields: Mapping[str, int], domain, h: float, device: int = 0):
        self._lib = abi.load()
        box = domain.boundarybox() if hasattr(domain, "boundarybox") else domain
        lo = (C.c_double * 3)(*[float(v) for v in box.lo])
        hi = (C.c_double * 3)(*[float(v) for v in box.hi])
        handle = C.c_void_p()
        abi.check(self._lib.sp_create(C.byref(handle), lo, hi, float(h), int(device)), None)
        self._h = handle
        self.h = float(h)
        self.domain = box
        self.device = device
        self.fields: Dict[str, int] = {"x": 3}
        self._fid: Dict[str, int] = {"x": 0}
        for name, nc in particle_fields.items():
            self.add_field(name, nc)
        phase = (C.c_int64 * 3)()
        lim = (C.c_int64 * 3)()
        kmax = C.c_int64()
        nd = C.c_int32()
        diff = (C.c_int64 * 27)()
        abi.check(self._lib.sp_key_params(self._h, phase, lim, C.byref(kmax), C.byref(nd), diff), self._h)
        self.key_phase = tuple(phase)
        self.key_lim = tuple(lim)
        self.key_max = kmax.value
        self.key_diff = list(diff[: nd.value])

    # -- life cycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.sp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- fields
    def add_field(self, name: str, ncomp: int = 1) -> int:
        fid = C.c_int32()
        abi.check(self._lib.sp_add_field(self._h, name.encode(), int(ncomp), C.byref(fid)), self._h)
        self.fields[name] = int(ncomp)
        self._fid[name] = fid.value
        return fid.value

    def fid(self, name: str) -> int:
        return self._fid[name]

    def __len__(self) -> int:
        n = C.c_int64()
        abi.check(self._lib.sp_num_particles(self._h, C.byref(n)), self._h)
        return n.value

    @property
    def n(self) -> int:
        return len(self)

    def resize(self, n: int):
        abi.check(self._lib.sp_resize(self._h, int(n)), self._h)

    def add_particles(self, **arrays):
        """``generate_particles!`` / ``push!`` (src/grids.jl:253-258): append particles at the end of the
        reference order.  ``x`` is required; missing fields are zero."""
        x = _farr(arrays["x"])
        if x.ndim != 2 or x.shape[1] != 3:
            raise ValueError("x must have shape (n, 3)")
        n_new = x.shape[0]
        n_old = len(self)
        if n_new == 0:
            return
        self.resize(n_old + n_new)
        for name, values in arrays.items():
            nc = self.fields[name]
            add = np.broadcast_to(_farr(values), (n_new,) if nc == 1 else (n_new, nc))
            if n_old:
                full = np.concatenate([self.get(name)[:n_old], add], axis=0)
            else:
                full = add
            self.set(name, full)   # fields not given stay zero (sp_resize zero-fills the new tail)

    def generate_particles(self, grid, shape, **constants) -> int:
        """``generate_particles!(sys, grid, shape, constructor)`` on the device (src/grids.jl:253-258): the lattice
        points of ``grid`` inside ``shape`` are appended in the reference's order with identical coordinates;
        ``constants`` are the fields the constructor sets to a constant (e.g. ``type=1.0, rho=1000.0``), everything
        else is zero.  Returns the number of particles added."""
        from . import geometry as geo
        nodes, offsets = geo.compile_shape(shape)
        arr = (abi.ShapeNode * len(nodes))()
        for k, (kind, a, b, p) in enumerate(nodes):
            arr[k].kind, arr[k].a, arr[k].b = kind, a, b
            for i, v in enumerate(p):
                arr[k].p[i] = v
        gk = {geo.Squaregrid: K["SP_GRID_SQUARE"], geo.Hexagrid: K["SP_GRID_HEXAGONAL"],
              geo.CubicGrid: K["SP_GRID_CUBIC"]}[type(grid)]
        irange = np.asarray(geo.lattice_index_box(grid, shape), dtype=np.int64)
        names = [n for n in constants if n != "x"]
        ff = np.asarray([self._fid[n] for n in names], dtype=np.int32) if names else np.zeros(1, dtype=np.int32)
        fv = _farr([constants[n] for n in names]) if names else np.zeros(1)
        n_off = 0 if offsets is None else len(offsets)
        off = _farr(offsets).ravel() if n_off else np.zeros(3)
        added = C.c_int64()
        abi.check(self._lib.sp_generate_particles(self._h, gk, float(grid.dr), arr, len(nodes), abi.ptr_f64(off), n_off,
                                                  abi.ptr_i64(irange), abi.ptr_i32(ff), abi.ptr_f64(fv), len(names),
                                                  C.byref(added)), self._h)
        return added.value

    def respawn(self, type_field: str, from_type: float, to_type: float, x1_min: float, shift: float,
                **constants) -> int:
        """``add_new_particles!`` of examples/cylinder.jl:145-156 on the device: particles of ``from_type`` with
        ``x[1] >= x1_min`` become ``to_type`` and a fresh ``from_type`` particle is appended ``shift`` upstream of
        each, in the order of their sources; ``constants`` are the fields the script's constructor sets
        (``rho=rho0, m=m0``), everything else is zero.  Returns the number of particles added."""
        names = list(constants)
        ff = np.asarray([self._fid[n] for n in names], dtype=np.int32) if names else np.zeros(1, dtype=np.int32)
        fv = _farr([constants[n] for n in names]) if names else np.zeros(1)
        added = C.c_int64()
        abi.check(self._lib.sp_respawn(self._h, self._fid[type_field], float(from_type), float(to_type), float(x1_min),
                                       float(shift), abi.ptr_i32(ff), abi.ptr_f64(fv), len(names), C.byref(added)),
                  self._h)
        return added.value

    def set(self, name: str, values):
        """Upload a field in reference order: shape (n,) or (n, ncomp)."""
        nc = self.fields[name]
        a = _farr(values)
        n = len(self)
        if a.size != n * nc:
            raise ValueError(f"field {name}: expected {n}x{nc} values, got shape {a.shape}")
        abi.check(self._lib.sp_upload(self._h, self._fid[name], abi.ptr_f64(a), n, K["SP_LAYOUT_AOS"]), self._h)

    def get(self, name: str) -> np.ndarray:
        """Download a field in reference order."""
        nc = self.fields[name]
        n = len(self)
        out = np.empty((n, nc) if nc > 1 else (n,), dtype=np.float64)
        if n:
            abi.check(self._lib.sp_download(self._h, self._fid[name], abi.ptr_f64(out), n, K["SP_LAYOUT_AOS"]),
                      self._h)
        return out

    def upload_raw(self, name: str, host_ptr, n: int, layout: int):
        abi.check(self._lib.sp_upload(self._h, self._fid[name], host_ptr, n, layout), self._h)

    def download_raw(self, name: str, host_ptr, n: int, layout: int):
        abi.check(self._lib.sp_download(self._h, self._fid[name], host_ptr, n, layout), self._h)

    def synchronize(self):
        abi.check(self._lib.sp_synchronize(self._h), self._h)

    # -- the hot path
    def create_cell_list(self):
        """``create_cell_list!(sys)`` (src/core.jl:51-90)."""
        abi.check(self._lib.sp_create_cell_list(self._h), self._h)

    def _bind(self, names: Sequence[str]):
        return np.asarray([self._fid[nm] for nm in names], dtype=np.int32)

    def apply(self, op: Operator, self_: bool = False, strict_order: bool = False, tile_kernel: bool = False,
              unfused_build: bool = False):
        """``apply!(sys, action!; self=false)`` (src/core.jl:151-161) for a registered operator."""
        F = self._bind(op.fields)
        P = _farr(op.params)
        flags = ((K["SP_FLAG_SELF"] if self_ else 0) | (K["SP_FLAG_STRICT_ORDER"] if strict_order else 0)
                 | (K["SP_FLAG_TILE_KERNEL"] if tile_kernel else 0)
                 | (K["SP_FLAG_UNFUSED_BUILD"] if unfused_build else 0))
        abi.check(self._lib.sp_apply(self._h, op.op, abi.ptr_i32(F), len(F), abi.ptr_f64(P), len(P), flags), self._h)

    def sum_at_points(self, sum_op: int, fields: Sequence[str], params: Sequence[float], points) -> np.ndarray:
        """``SmoothedParticles.sum(sys, f, x)`` (src/core.jl:240-260) for many points at once."""
        pts = _farr(points).reshape(-1, 3)
        out = np.empty(len(pts))
        F = self._bind(fields)
        P = _farr(params)
        abi.check(self._lib.sp_sum_at_points(self._h, sum_op, abi.ptr_i32(F), len(F), abi.ptr_f64(P), len(P),
                                             abi.ptr_f64(pts), len(pts), abi.ptr_f64(out)), self._h)
        return out

    def reduce(self, red: int, fields: Sequence[str], params: Sequence[float] = (), nout: int = 1) -> np.ndarray:
        F = self._bind(fields)
        P = _farr(params) if len(params) else np.zeros(1)
        out = np.zeros(3)
        abi.check(self._lib.sp_reduce(self._h, red, abi.ptr_i32(F), len(F), abi.ptr_f64(P), len(params),
                                      abi.ptr_f64(out)), self._h)
        return out[:nout].copy()

    def assemble_vector(self, op: Operator) -> np.ndarray:
        """``assemble_vector(sys, func)`` (src/core.jl:175-182): the unary operator writes its LAST bound
        field, which is then downloaded."""
        self.apply(op)
        return self.get(op.fields[-1])

    def poisson_apply(self, A: PoissonOperator, p_in: str, y_out: str):
        F = self._bind(tuple(A.fields) + (p_in, y_out))
        P = _farr(A.params)
        abi.check(self._lib.sp_poisson_apply(self._h, abi.ptr_i32(F), len(F), abi.ptr_f64(P), len(P)), self._h)

    def assemble_matrix(self, A: PoissonOperator):
        """``assemble_matrix(sys, projection_matrix)`` (src/core.jl:196-225) as COO triplets (I, J, V): 1-based
        reference indices, the diagonal included, duplicates (narrow-domain double visits) left for the caller's
        sparse constructor to sum — what the reference passes to ``sparse(I, J, V, N, N)``.  Evaluated on the
        device (ELL coefficients on the cached neighbour lists) and copied out."""
        F = self._bind(tuple(A.fields))
        P = _farr(A.params)
        nnz = C.c_int64()
        abi.check(self._lib.sp_assemble_matrix(self._h, abi.ptr_i32(F), len(F), abi.ptr_f64(P), len(P), None, None, None,
                                               0, C.byref(nnz)), self._h)
        I = np.empty(max(nnz.value, 1), dtype=np.int64)
        J = np.empty(max(nnz.value, 1), dtype=np.int64)
        V = np.empty(max(nnz.value, 1))
        if nnz.value:
            abi.check(self._lib.sp_assemble_matrix(self._h, abi.ptr_i32(F), len(F), abi.ptr_f64(P), len(P), abi.ptr_i64(I),
                                                   abi.ptr_i64(J), abi.ptr_f64(V), nnz.value, C.byref(nnz)), self._h)
        return I[:nnz.value], J[:nnz.value], V[:nnz.value]

    def poisson_cg(self, A: PoissonOperator, b: str, P_out: str, reltol=None, abstol=0.0, maxiter=0):
        """``P .= cg(A, b)`` (collapse_dry_implicit.jl:227) without assembling A."""
        if reltol is None:
            reltol = float(np.sqrt(np.finfo(np.float64).eps))
        F = self._bind(tuple(A.fields) + (b, P_out))
        P = _farr(A.params)
        iters = C.c_int64()
        resid = C.c_double()
        rc = self._lib.sp_poisson_cg(self._h, abi.ptr_i32(F), len(F), abi.ptr_f64(P), len(P), float(reltol),
                                     float(abstol), int(maxiter), C.byref(iters), C.byref(resid))
        # IterativeSolvers.cg hands back its last iterate when maxiter is reached, without an error, and the script keeps
        # stepping (collapse_dry_implicit.jl:223-227): SP_ERR_NOT_CONVERGED is a soft status here, P_out holds that
        # iterate, and the caller reads `last_cg_converged`
        self.last_cg_converged = rc != K["SP_ERR_NOT_CONVERGED"]
        if rc != K["SP_ERR_NOT_CONVERGED"]:
            abi.check(rc, self._h)
        return iters.value, resid.value

    def run_program(self, program: int, fields: Sequence[str], params: Sequence[float], nsteps: int):
        F = self._bind(fields)
        P = _farr(params)
        abi.check(self._lib.sp_run_program(self._h, program, abi.ptr_i32(F), len(F), abi.ptr_f64(P), len(P),
                                           int(nsteps)), self._h)

    # -- step graphs: a loop body recorded once, replayed as one CUDA graph launch (sp_graph_* of include/sp_b200.h)
    def record(self, body, repeat: int = 2) -> "StepGraph":
        """Run ``body()`` ``repeat`` times as ONE recorded unit and return a replayable graph.

        ``body`` is the host's loop body (``lambda: case.step(sys)``): apply / create_cell_list calls only.  The unit is
        executed once by this call (like the calls it encloses).  ``repeat`` must make the number of cell-list builds in
        the unit even (2 time steps when a step builds once).  Run the body the ordinary way at least once before."""
        abi.check(self._lib.sp_graph_begin(self._h), self._h)
        try:
            for _ in range(repeat):
                body()
        except BaseException:
            gid = C.c_int32()
            self._lib.sp_graph_end(self._h, C.byref(gid))   # close the recording; its status is secondary here
            raise
        gid = C.c_int32()
        abi.check(self._lib.sp_graph_end(self._h, C.byref(gid)), self._h)
        return StepGraph(self, gid.value, repeat)

    # -- parity / debug views (1-based, reference order)
    def cell_keys(self) -> np.ndarray:
        n = len(self)
        out = np.empty(n, dtype=np.int64)
        if n:
            abi.check(self._lib.sp_get_cell_keys(self._h, abi.ptr_i64(out), n), self._h)
        return out

    def cell_list(self):
        n = len(self)
        offsets = np.empty(self.key_max + 1, dtype=np.int64)
        members = np.empty(max(n, 1), dtype=np.int64)
        abi.check(self._lib.sp_get_cell_list(self._h, abi.ptr_i64(offsets), abi.ptr_i64(members)), self._h)
        return offsets, members[:n]

    def neighbour_lists(self):
        n = len(self)
        offsets = np.zeros(n + 1, dtype=np.int64)
        abi.check(self._lib.sp_get_neighbour_lists(self._h, abi.ptr_i64(offsets), None, 0), self._h)
        total = int(offsets[n])
        ids = np.empty(max(total, 1), dtype=np.int64)
        abi.check(self._lib.sp_get_neighbour_lists(self._h, abi.ptr_i64(offsets), abi.ptr_i64(ids), total), self._h)
        return offsets, ids[:total]

    def build_neighbour_lists(self):
        """Build the cached neighbour lists now (optional: the first binary apply() after a move does it lazily)."""
        abi.check(self._lib.sp_build_neighbour_lists(self._h), self._h)

    @property
    def neighbour_list_capacity(self) -> int:
        k = C.c_int32()
        abi.check(self._lib.sp_neighbour_list_capacity(self._h, C.byref(k)), self._h)
        return k.value

    def sweep_neighbour_lists(self):
        """The cached lists the default pair sweeps replay (same sets as neighbour_lists(), sweep visiting order)."""
        n = len(self)
        offsets = np.zeros(n + 1, dtype=np.int64)
        abi.check(self._lib.sp_get_sweep_neighbour_lists(self._h, abi.ptr_i64(offsets), None, 0), self._h)
        total = int(offsets[n])
        ids = np.empty(max(total, 1), dtype=np.int64)
        abi.check(self._lib.sp_get_sweep_neighbour_lists(self._h, abi.ptr_i64(offsets), abi.ptr_i64(ids), total), self._h)
        return offsets, ids[:total]

    @property
    def n_removed(self) -> int:
        n = C.c_int64()
        abi.check(self._lib.sp_num_removed(self._h, C.byref(n)), self._h)
        return n.value

    def last_call_ms(self) -> float:
        ms = C.c_float()
        abi.check(self._lib.sp_last_call_ms(self._h, C.byref(ms)), self._h)
        return ms.value

    def timer_start(self):
        abi.check(self._lib.sp_timer_start(self._h), self._h)

    def timer_stop(self) -> float:
        ms = C.c_float()
        abi.check(self._lib.sp_timer_stop(self._h, C.byref(ms)), self._h)
        return ms.value

    @property
    def launch_count(self) -> int:
        n = C.c_int64()
        abi.check(self._lib.sp_launch_count(self._h, C.byref(n)), self._h)
        return n.value


# free functions with the reference's names (src/SmoothedParticles.jl:10-72)
def create_cell_list(sys: ParticleSystem):
    sys.create_cell_list()


def cfl_time_step(sys, cfl: float, h: float, c: float, v: str = "v") -> float:
    """Adaptive time step dt = cfl*h/(c + max|v|) (BASELINE north_star: "allreduce for the CFL time-step minimum").
    One device reduction (SP_RED_MAX_SPEED) and an 8-byte read-back; on a slab system the maximum is all-reduced over
    the ranks, so every rank computes the same dt.  The reference's examples use fixed steps (dt = 0.1*h/c): this is
    an extension with no upstream counterpart."""
    vmax = float(sys.reduce(K["SP_RED_MAX_SPEED"], (v,), (), nout=1)[0])
    return cfl * h / (c + vmax)


def apply(sys: ParticleSystem, op: Operator, self: bool = False, strict_order: bool = False):
    sys.apply(op, self_=self, strict_order=strict_order)


def apply_unary(sys: ParticleSystem, op: Operator):
    """``apply_unary!(sys, action!)`` (src/core.jl:138-142): ``op`` must be a per-particle operator."""
    if op.binary:
        raise TypeError(f"{op.name or op.op}: a binary operator was passed to apply_unary")
    sys.apply(op)


def apply_binary(sys: ParticleSystem, op: Operator, strict_order: bool = False):
    """``apply_binary!(sys, action!)`` (src/core.jl:125-129): the pair sweep without the ``self`` term."""
    if not op.binary:
        raise TypeError(f"{op.name or op.op}: a unary operator was passed to apply_binary")
    sys.apply(op, self_=False, strict_order=strict_order)


def assemble_vector(sys: ParticleSystem, op: Operator) -> np.ndarray:
    return sys.assemble_vector(op)


def assemble_matrix(sys: ParticleSystem, A: PoissonOperator):
    """``assemble_matrix(sys, func)`` (src/core.jl:196-225): the N x N matrix as ``scipy.sparse.csc_matrix`` (the
    reference returns a SparseMatrixCSC); duplicate triplets are summed, as ``sparse()`` does."""
    import scipy.sparse as sps
    I, J, V = sys.assemble_matrix(A)
    n = len(sys)
    return sps.coo_matrix((V, (I - 1, J - 1)), shape=(n, n)).tocsc()


def _named_kernel(kernel: str, kfun: str):
    def f(h: float, r, device: int = 0):
        """Kernel function of src/kernels.jl evaluated on the device (scalar or array ``r``)."""
        out = kernel_eval(kernel, K[kfun], h, np.atleast_1d(_farr(r)), device)
        return float(out[0]) if np.ndim(r) == 0 else out
    return f


# the kernel functions under the reference's names (src/SmoothedParticles.jl:23-27): wendland2(h, r), rDspline23(h, r), ...
KERNEL_FUNCTIONS = {}
for _k in ("wendland1", "wendland2", "wendland3", "spline23", "spline24"):
    KERNEL_FUNCTIONS[_k] = _named_kernel(_k, "SP_KFUN_W")
    KERNEL_FUNCTIONS["D" + _k] = _named_kernel(_k, "SP_KFUN_DW")
    KERNEL_FUNCTIONS["rD" + _k] = _named_kernel(_k, "SP_KFUN_RDW")
KERNEL_FUNCTIONS["DDwendland3"] = _named_kernel("wendland3", "SP_KFUN_DDW")
globals().update(KERNEL_FUNCTIONS)


class ParticleField:
    """``ParticleField(sys, :var)`` (src/structs.jl:118-125): array-like view of one scalar field.
    Reads download, ``field[:] = values`` uploads."""

    def __init__(self, sys: ParticleSystem, name: str):
        self.sys = sys
        self.name = name

    def __len__(self):
        return len(self.sys)

    def __array__(self, dtype=None, copy=None):
        return self.sys.get(self.name)

    def __getitem__(self, idx):
        return self.sys.get(self.name)[idx]

    def __setitem__(self, idx, val):
        cur = self.sys.get(self.name)
        cur[idx] = val
        self.sys.set(self.name, cur)


def kernel_eval(kernel, kfun: int, h: float, r, device: int = 0) -> np.ndarray:
    """Evaluate a kernel function of src/kernels.jl on the device."""
    lib = abi.load()
    r = _farr(r)
    out = np.empty_like(r)
    kid = abi.KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel)
    abi.check(lib.sp_kernel_eval(kid, int(kfun), float(h), abi.ptr_f64(r), abi.ptr_f64(out), r.size, device), None)
    return out


class StepGraph:
    """Handle of a recorded loop body (``ParticleSystem.record``): ``replay(k)`` runs the unit k times, one CUDA graph
    launch each; ``steps_per_replay`` says how many executions of the body one replay is."""

    def __init__(self, system: "ParticleSystem", gid: int, steps_per_replay: int):
        self._sys, self._gid, self.steps_per_replay = system, gid, steps_per_replay

    def replay(self, times: int = 1):
        abi.check(self._sys._lib.sp_graph_launch(self._sys._h, self._gid, int(times)), self._sys._h)

    def close(self):
        if self._gid >= 0 and self._sys._h:
            self._sys._lib.sp_graph_destroy(self._sys._h, self._gid)
        self._gid = -1
