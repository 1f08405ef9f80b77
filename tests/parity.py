"""Helpers shared by the GPU parity tests: run the same program on the CUDA library (through the C ABI)
and on the CPU oracle, and compare in reference order."""
import numpy as np

# north_star tolerance: per-step forces, density, pressure within 1e-10 relative (FP64, different
# summation order); neighbour lists and cell assignment bit-exact.
RTOL_STEP = 1e-10


def rel_err(a, b, floor=0.0):
    """Norm-wise relative error of a against b (both arrays), with an absolute floor on the scale."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, floor)
    if scale == 0.0:
        return float(np.max(np.abs(a - b))) if a.size else 0.0
    return float(np.max(np.abs(a - b))) / scale


# Element-wise bar: |a_i - b_i| <= rtol * max(|b_i|, floor), floor = the larger of the field's absolute
# floor (if the test gives one) and ELEM_FLOOR_FRAC * max|b|.  A pair sum that cancels to a small value carries the rounding error of
# its LARGE terms (the hydrostatic pressure terms of Dv cancel to ~1e-3 of their size), so a pure element-wise relative
# error is not meaningful below that scale; above it every element is held to the same rtol as the max-norm.
ELEM_FLOOR_FRAC = 1e-2


def elem_err(a, b, floor=0.0):
    """max_i |a_i - b_i| / max(|b_i|, floor_abs), floor_abs = max(floor, ELEM_FLOOR_FRAC * max|b|)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    fin = np.isfinite(b)
    if not np.any(fin):
        return 0.0
    bmax = float(np.max(np.abs(b[fin])))
    fl = max(floor, ELEM_FLOOR_FRAC * bmax)
    if fl == 0.0:
        return float(np.max(np.abs(a[fin] - b[fin])))
    den = np.maximum(np.abs(b[fin]), fl)
    return float(np.max(np.abs(a[fin] - b[fin]) / den))


def assert_fields_close(dev, ora, names, rtol=RTOL_STEP, what="", floors=None, etol=None):
    """Two bars per field, both at `rtol`: the norm-wise error max|a-b| / max|b| and the element-wise error of
    ``elem_err``.  floors: per-field absolute scale below which a field is rounding noise (e.g. a pressure that is the
    difference of two equal densities): both errors are then measured against that physical scale."""
    assert len(dev) == len(ora), f"{what}: particle counts differ {len(dev)} vs {len(ora)}"
    floors = floors or {}
    for nm in names:
        a, b = dev.get(nm), ora.get(nm)
        assert np.all(np.isfinite(a) == np.isfinite(b)), f"{what}: field {nm} finiteness differs"
        e = rel_err(a, b, floors.get(nm, 0.0))
        assert e <= rtol, f"{what}: field {nm} relative error {e:.3e} > {rtol:.1e}"
        # element-wise at the north-star bar (1e-10) even where the norm-wise bar of a test is tighter: a field like
        # P = c^2 (rho - rho0) is a difference of nearly equal numbers, its small elements carry the absolute rounding
        # error of rho (1e-16 * rho0 * c^2), which is ~1e-13 of the field's own maximum
        et = max(rtol, RTOL_STEP) if etol is None else etol
        ee = elem_err(a, b, floors.get(nm, 0.0))
        assert ee <= et, f"{what}: field {nm} element-wise error {ee:.3e} > {et:.1e}"


def neighbour_sets_equal(dev, ora, ordered=False):
    od, idd = dev.neighbour_lists()
    oo, ido = ora.neighbour_lists()
    if not np.array_equal(od, oo):
        return False
    if ordered:
        return np.array_equal(idd, ido)
    # compare as sets per particle: sort each segment
    seg = np.repeat(np.arange(len(od) - 1), np.diff(od))
    a = np.lexsort((idd, seg))
    b = np.lexsort((ido, seg))
    return np.array_equal(idd[a], ido[b])
