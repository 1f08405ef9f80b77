"""Host side of the slab decomposition (one process per GPU): which rank owns which cell layers, how the
NCCL unique id reaches every rank, and the slab-aware time-step drivers.  The exchange itself (migration,
ghost halos, all-reduces) happens inside the CUDA library over NCCL — see csrc/sp_slab.cu."""
from __future__ import annotations

import ctypes as C
import math
from typing import Sequence, Tuple

import numpy as np

from . import abi
from .system import ParticleSystem

K = abi.K


def slab_axis(key_lim: Sequence[int]) -> int:
    """The slowest axis of the linear cell key: z in 3-D, y in 2-D (key_lim[2] == 1, src/structs.jl:70)."""
    return 1 if key_lim[2] == 1 else 2


def partition_layers(n_layers: int, nranks: int) -> list:
    """Owned cell layers [c0, c1) of every rank: nearly equal counts, the first `rem` ranks get one more."""
    base, rem = divmod(n_layers, nranks)
    out, c = [], 0
    for r in range(nranks):
        w = base + (1 if r < rem else 0)
        out.append((c, c + w))
        c += w
    return out


def balanced_cuts(layer_counts: Sequence[int], nranks: int, min_layers: int = 3) -> list:
    """Cut planes that balance the particle counts: ``layer_counts[l]`` particles in cell layer l -> nranks + 1 layer
    numbers, every slab at least ``min_layers`` layers thick (sp_slab_init_cuts).  Greedy on the cumulative count: the
    r-th cut goes where the running total is closest to r/nranks of all particles."""
    cnt = np.asarray(layer_counts, dtype=np.float64)
    nl = len(cnt)
    assert nl >= min_layers * nranks, "too few cell layers for this many ranks"
    cum = np.concatenate([[0.0], np.cumsum(cnt)])
    cuts = [0]
    for r in range(1, nranks):
        target = cum[-1] * r / nranks
        lo = cuts[-1] + min_layers
        hi = nl - min_layers * (nranks - r)
        c = int(np.argmin(np.abs(cum[lo:hi + 1] - target))) + lo
        cuts.append(c)
    cuts.append(nl)
    return cuts


def owner_of(x_axis: np.ndarray, h: float, key_phase_axis: int, layers: list) -> np.ndarray:
    """Rank owning each coordinate along the slab axis (same floor(x/h) as find_key, src/structs.jl:99-101).
    Returns -1 for coordinates outside every slab."""
    cell = np.floor(np.asarray(x_axis) / h).astype(np.int64) - key_phase_axis
    owner = np.full(cell.shape, -1, dtype=np.int64)
    for r, (c0, c1) in enumerate(layers):
        owner[(cell >= c0) & (cell < c1)] = r
    return owner


def unique_id() -> bytes:
    lib = abi.load()
    buf = (C.c_uint8 * 128)()
    abi.check(lib.sp_slab_unique_id(buf), None)
    return bytes(buf)


class SlabSystem(ParticleSystem):
    """A ParticleSystem cut into slabs over `nranks` processes.  Every rank constructs it with the same global
    domain and h, then adds only the particles it owns (``owner_of``)."""

    def __init__(self, particle_fields, domain, h, rank: int, nranks: int, nccl_id: bytes, periodic: bool = False,
                 device: int = 0, cuts: Sequence[int] | None = None):
        """cuts: optional ``nranks + 1`` cell-layer numbers (0 .. key_lim[axis]); rank r owns layers [cuts[r], cuts[r+1]).
        Default: equal layer counts (``partition_layers``); ``balanced_cuts`` balances particle counts instead."""
        super().__init__(particle_fields, domain, h, device=device)
        self.global_key_phase = self.key_phase
        self.global_key_lim = self.key_lim
        self.rank, self.nranks, self.periodic = rank, nranks, periodic
        idbuf = (C.c_uint8 * 128).from_buffer_copy(nccl_id)
        if cuts is None:
            abi.check(self._lib.sp_slab_init(self._h, idbuf, rank, nranks, 1 if periodic else 0), self._h)
        else:
            carr = np.ascontiguousarray(cuts, dtype=np.int64)
            assert carr.size == nranks + 1
            abi.check(self._lib.sp_slab_init_cuts(self._h, idbuf, rank, nranks, 1 if periodic else 0, abi.ptr_i64(carr)),
                      self._h)
        for hidden in ("_ghost", "_gid"):
            fid = C.c_int32()
            abi.check(self._lib.sp_find_field(self._h, hidden.encode(), C.byref(fid)), self._h)
            self.fields[hidden] = 1
            self._fid[hidden] = fid.value
        c0, c1 = C.c_int64(), C.c_int64()
        lo, hi = C.c_double(), C.c_double()
        ax = C.c_int32()
        abi.check(self._lib.sp_slab_range(self._h, C.byref(c0), C.byref(c1), C.byref(lo), C.byref(hi), C.byref(ax)),
                  self._h)
        self.layers = (c0.value, c1.value)
        self.axis = ax.value
        self.coord_range = (lo.value, hi.value)
        # key parameters now describe the LOCAL window
        phase = (C.c_int64 * 3)()
        lim = (C.c_int64 * 3)()
        kmax = C.c_int64()
        nd = C.c_int32()
        diff = (C.c_int64 * 27)()
        abi.check(self._lib.sp_key_params(self._h, phase, lim, C.byref(kmax), C.byref(nd), diff), self._h)
        self.key_phase, self.key_lim, self.key_max = tuple(phase), tuple(lim), kmax.value
        self.key_diff = list(diff[: nd.value])

    def owns(self, x: np.ndarray) -> np.ndarray:
        cell = np.floor(np.asarray(x)[:, self.axis] / self.h).astype(np.int64) - self.global_key_phase[self.axis]
        return (cell >= self.layers[0]) & (cell < self.layers[1])

    def create_cell_list(self):
        """Migration + ghost halo + local build (sp_slab_create_cell_list)."""
        abi.check(self._lib.sp_slab_create_cell_list(self._h), self._h)

    def halo_refresh(self, *names: str):
        F = self._bind(names)
        abi.check(self._lib.sp_slab_halo_refresh(self._h, abi.ptr_i32(F), len(F)), self._h)

    @property
    def n_owned(self) -> int:
        n = C.c_int64()
        abi.check(self._lib.sp_slab_num_owned(self._h, C.byref(n)), self._h)
        return n.value

    def allreduce(self, values, op: str = "sum") -> np.ndarray:
        a = np.ascontiguousarray(values, dtype=np.float64).copy()
        abi.check(self._lib.sp_slab_allreduce(self._h, abi.ptr_f64(a), a.size, 1 if op == "max" else 0), self._h)
        return a

    def owned_mask(self) -> np.ndarray:
        return self.get("_ghost") == 0.0


def wcsph3d_slab_step(sys: SlabSystem, ops: dict):
    """examples/collapse3d.jl:136-150 on a slab system: the rebuild migrates particles and exchanges two ghost layers per
    side in one round; the inner ghost layer integrates its own density (it sees all of its neighbours), so no field has
    to be refreshed between the sweeps."""
    sys.apply(ops["move"])
    sys.create_cell_list()
    sys.apply(ops["bom"])
    sys.apply(ops["fp"])
    sys.apply(ops["force"])
    sys.apply(ops["acc"])
    sys.apply(ops["acc"])
