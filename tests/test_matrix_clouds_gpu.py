"""assemble_matrix export (src/core.jl:196-225) and the device cell list on randomized corner-case clouds, against the oracle."""
import numpy as np
import pytest

import smoothedparticles_jl_b200 as sp
from smoothedparticles_jl_b200 import ParticleSystem, configs
from oracle.oracle import OracleSystem

pytestmark = pytest.mark.gpu


def _presolve(case, s):
    o = case.ops
    case.prologue(s)
    s.apply(o["init"])
    s.create_cell_list()
    s.apply(o["visc"])
    s.apply(o["dll"])
    s.apply(o["b"])


def test_assemble_matrix_export_matches_the_oracle():
    # sp_assemble_matrix: assemble_matrix(sys, projection_matrix), src/core.jl:196-225, as COO triplets from the device
    import scipy.sparse as sps
    case = configs.collapse_dry_implicit(dr=2.0e-2)
    rng = np.random.default_rng(2)
    case.init["v"] = rng.uniform(-1, 1, size=(case.n, 3)) * np.array([1.0, 1.0, 0.0])
    dev, ora = case.make(ParticleSystem), case.make(OracleSystem)
    _presolve(case, dev)
    _presolve(case, ora)
    n = len(ora)
    Io, Jo, Vo = ora.assemble_matrix(case.ops["A"])
    Id, Jd, Vd = dev.assemble_matrix(case.ops["A"])
    assert len(Id) == len(Io)                                      # same triplet count: neighbours + diagonal
    Ao = sps.coo_matrix((Vo, (Io - 1, Jo - 1)), shape=(n, n)).tocsr()
    Ad = sps.coo_matrix((Vd, (Id - 1, Jd - 1)), shape=(n, n)).tocsr()
    Ao.sort_indices()
    Ad.sort_indices()
    assert np.array_equal(Ao.indptr, Ad.indptr) and np.array_equal(Ao.indices, Ad.indices)   # same sparsity pattern
    assert np.max(np.abs(Ao.data - Ad.data)) <= 1e-10 * np.max(np.abs(Ao.data))
    A = sp.assemble_matrix(dev, case.ops["A"])
    assert A.shape == (n, n) and abs(A - A.T).max() <= 1e-10 * np.max(np.abs(Ao.data))
    # the exported matrix and the matrix-free operator are the same operator
    p = rng.uniform(-1, 1, n)
    dev.set("P", p)
    dev.add_field("y", 1)
    dev.poisson_apply(case.ops["A"], "P", "y")
    assert np.max(np.abs(dev.get("y") - A @ p)) <= 1e-10 * np.max(np.abs(A @ p))
    # a second call after the particles moved and were re-sorted
    # (the oracle continues from the DEVICE state: after a CG solve to sqrt(eps) the two position sets differ by ~1e-8,
    # enough to flip a pair at r ~ h in or out of the pattern — bit-exactness is a statement about identical inputs)
    for s in (dev, ora):
        s.apply(case.ops["force"])
        s.apply(case.ops["acc"])
    for f in ("x", "v"):
        ora.set(f, dev.get(f))
    for s in (dev, ora):
        _presolve(case, s)
    Io, Jo, Vo = ora.assemble_matrix(case.ops["A"])
    Id, Jd, Vd = dev.assemble_matrix(case.ops["A"])
    Ao = sps.coo_matrix((Vo, (Io - 1, Jo - 1)), shape=(n, n)).tocsr()
    Ad = sps.coo_matrix((Vd, (Id - 1, Jd - 1)), shape=(n, n)).tocsr()
    Ao.sort_indices()
    Ad.sort_indices()
    assert np.array_equal(Ao.indptr, Ad.indptr) and np.array_equal(Ao.indices, Ad.indices)
    assert abs(Ao - Ad).max() <= 1e-9 * np.max(np.abs(Ao.data))


@pytest.mark.parametrize("dim", [2, 3])
def test_device_cell_list_on_random_clouds(dim):
    # the randomized corner cases of test_oracle_against_a_literal_python_port_on_random_inputs, device vs oracle:
    # survivors and numbering, per-cell member lists, neighbour lists in visiting order — all bit-exact
    from smoothedparticles_jl_b200 import geometry as geo
    from test_oracle_pins import _random_cloud
    rng = np.random.default_rng(200 + dim)
    for trial in range(40):
        h, lo, hi, x = _random_cloud(rng, dim)
        n = len(x)
        box = geo.Box(*lo, *hi)
        dev, ora = ParticleSystem({"tag": 1}, box, h), OracleSystem({"tag": 1}, box, h)
        assert tuple(dev.key_lim) == tuple(ora.key_lim) and dev.key_max == ora.key_max
        if n:
            for s in (dev, ora):
                s.add_particles(x=x, tag=np.arange(1, n + 1, dtype=float))
        for rebuild in range(2):
            dev.create_cell_list()
            ora.create_cell_list()
            assert len(dev) == len(ora)
            m = len(ora)
            if m == 0:
                break
            assert np.array_equal(dev.get("tag"), ora.get("tag"))
            assert np.array_equal(dev.cell_keys(), ora.cell_keys())
            (od, md), (oo, mo) = dev.cell_list(), ora.cell_list()
            assert np.array_equal(od, oo) and np.array_equal(md, mo)
            (nd, idd), (no, ido) = dev.neighbour_lists(), ora.neighbour_lists()
            assert np.array_equal(nd, no) and np.array_equal(idd, ido)
            (sd, sid), _ = dev.sweep_neighbour_lists(), None
            assert np.array_equal(np.diff(sd), np.diff(no))          # the cached lists hold the same sets
            xs = ora.get("x") + rng.uniform(-0.3, 0.3, (m, 3)) * h * (1.0 if dim == 3 else np.array([1.0, 1.0, 0.0]))
            dev.set("x", xs)
            ora.set("x", xs)


def test_assemble_matrix_on_a_narrow_domain_counts_the_diagonal_per_visit():
    # core.jl:196-225 does not skip p == q: on a domain with fewer than three cells along an axis the linear key offsets
    # visit a particle's own cell more than once, the diagonal triplet is pushed once per visit and sparse() sums them.
    # Export, matrix-free operator and CG operator must all be that matrix.
    import scipy.sparse as sps
    from smoothedparticles_jl_b200 import geometry as geo, operators as ops
    rng = np.random.default_rng(12)
    h = 0.1
    box = geo.Box(0.0, 0.0, 0.0, 0.05, 1.0, 0.0)            # 1 x 11 x 1 cells: offsets di + 1*dj, three of them zero
    n = 70                                                  # ~14 distinct neighbours, each visited three times
    x = np.column_stack([rng.uniform(0.0, 0.05, n), rng.uniform(0.0, 1.0, n), np.zeros(n)])
    fields = {"L": 1, "lambda": 1, "type": 1, "P": 1, "y": 1}
    vals = dict(L=rng.uniform(0.5, 2.0, n), type=(rng.uniform(size=n) < 0.3).astype(float))
    vals["lambda"] = rng.uniform(-1.0, 1.0, n)
    dev, ora = ParticleSystem(fields, box, h), OracleSystem(fields, box, h)
    for s in (dev, ora):
        s.add_particles(x=x, **vals)
        s.create_cell_list()
    assert sum(1 for d in ora.key_diff if d == 0) > 1, "the test needs a domain with repeated visits of the own cell"
    A = ops.isph_projection_matrix("spline23", 1.0e-3, h, 1.0, 0.7)
    Io, Jo, Vo = ora.assemble_matrix(A)
    Id, Jd, Vd = dev.assemble_matrix(A)
    assert len(Id) == len(Io)
    Ao = sps.coo_matrix((Vo, (Io - 1, Jo - 1)), shape=(n, n)).tocsr()
    Ad = sps.coo_matrix((Vd, (Id - 1, Jd - 1)), shape=(n, n)).tocsr()
    assert abs(Ao - Ad).max() <= 1e-10 * np.max(np.abs(Ao.data))
    p = rng.uniform(-1, 1, n)
    dev.set("P", p)
    dev.poisson_apply(A, "P", "y")
    assert np.max(np.abs(dev.get("y") - Ao @ p)) <= 1e-10 * np.max(np.abs(Ao @ p))
