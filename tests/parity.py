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


def assert_fields_close(dev, ora, names, rtol=RTOL_STEP, what="", floors=None):
    """floors: per-field absolute scale below which a field is rounding noise (e.g. a pressure that is the
    difference of two equal densities): the error is then measured against that physical scale."""
    assert len(dev) == len(ora), f"{what}: particle counts differ {len(dev)} vs {len(ora)}"
    floors = floors or {}
    for nm in names:
        a, b = dev.get(nm), ora.get(nm)
        assert np.all(np.isfinite(a) == np.isfinite(b)), f"{what}: field {nm} finiteness differs"
        e = rel_err(a, b, floors.get(nm, 0.0))
        assert e <= rtol, f"{what}: field {nm} relative error {e:.3e} > {rtol:.1e}"


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
