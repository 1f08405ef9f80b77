"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header declares,
argument validation works without a device, and the host-side set-up code reproduces the configs."""
import ctypes as C
import os

import numpy as np
import pytest

import smoothedparticles_jl_b200 as sp
from smoothedparticles_jl_b200 import configs, geometry as geo

K = sp.K


def test_library_exports_every_declared_symbol():
    lib = sp.abi.load()
    declared = sp.abi.declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sp_b200.h but not exported"
        assert name in sp.abi.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.sp_version() == K["SP_ABI_VERSION"]


def test_argument_validation_without_device():
    lib = sp.abi.load()
    h = C.c_void_p()
    lo = (C.c_double * 3)(0, 0, 0)
    hi = (C.c_double * 3)(1, 1, 1)
    # structs.jl:59 — h must be positive
    assert lib.sp_create(C.byref(h), lo, hi, 0.0, 0) == K["SP_ERR_INVALID"]
    assert b"h must be a positive float" in lib.sp_last_error(None)
    assert lib.sp_create(None, lo, hi, 0.1, 0) == K["SP_ERR_INVALID"]
    n = C.c_int32(-1)
    rc = lib.sp_device_count(C.byref(n))
    if rc != 0 or n.value == 0:
        # no GPU here: creation must fail loudly, never fall back to a CPU path
        assert lib.sp_create(C.byref(h), lo, hi, 0.1, 0) == K["SP_ERR_NO_DEVICE"]
        with pytest.raises(sp.SpError):
            sp.ParticleSystem({}, geo.Box(0, 0, 0, 1, 1, 1), 0.1)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "smoothedparticles.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "libsp_oracle", "sp_oracle.cpp", "so_create"):
                    assert bad not in text, f"{f} references the oracle ({bad})"


def test_config_particle_counts():
    # counts of the shipped configs
    assert configs.collapse_dry().n == 10363
    c = configs.cavity_flow()
    assert c.n == 11219 and int((c.init["type"] == 2.0).sum()) == 289
    assert configs.collapse_dry_implicit().n == 23172
    c3 = configs.collapse3d()
    assert c3.n == 103479 and int((c3.init["type"] == 0.0).sum()) == 53041
    assert configs.collision_2d().n == 2510


def test_covering_order_and_hexagrid():
    g = geo.Squaregrid(0.5)
    X = geo.covering(g, geo.Rectangle(0.0, 0.0, 1.0, 1.0))
    # i outer, j inner (src/grids.jl:58)
    assert np.array_equal(X[:, :2], np.array([[i * 0.5, j * 0.5] for i in range(3) for j in range(3)]))
    hgrid = geo.Hexagrid(1.0)
    assert hgrid.a == pytest.approx((4 / 3) ** 0.25) and hgrid.b == pytest.approx((3 / 4) ** 0.25)
    X = geo.covering(hgrid, geo.Rectangle(-2.0, -2.0, 2.0, 2.0))
    # odd rows are shifted by half a cell with the sign of j (C remainder, src/grids.jl:83)
    rows = np.round(X[:, 1] / hgrid.b).astype(int)
    frac = np.round(X[:, 0] / hgrid.a * 2).astype(int) % 2
    assert np.array_equal(frac, np.abs(rows) % 2)


def test_boundary_layer_matches_pointwise_definition():
    grid = geo.Squaregrid(0.1)
    box = geo.Rectangle(0.0, 0.0, 1.0, 1.0)
    bl = geo.BoundaryLayer(box, grid, 0.25)
    pts = geo.covering(grid, bl)
    assert len(pts) > 0 and not np.any(box.is_inside(pts))
    # brute force: geometry.jl:208-218
    allp = geo.covering(grid, geo.Rectangle(-0.5, -0.5, 1.5, 1.5))
    expect = []
    for x in allp:
        if box.is_inside(x[None])[0]:
            continue
        if any(box.is_inside((x + dx)[None])[0] for dx in bl.dxs):
            expect.append(x)
    assert np.array_equal(pts, np.array(expect))


def test_operator_parameter_folding():
    from smoothedparticles_jl_b200 import operators as ops
    o = ops.balance_of_mass("wendland3", 2.0, 0.1, 1e-4)
    assert o.op == K["SP_OP_BALANCE_OF_MASS"] and o.params == (3.0, 2.0, 0.1, 2 * 1e-4) and o.binary
    assert ops.find_pressure(1e-3, 50.0, 1000.0).params == (1e-3, 2500.0, 1000.0, 0.0)
    assert ops.fill("a").fields == ("a",)


def test_reference_surface_names_exist():
    # the exports of src/SmoothedParticles.jl:10-72 that belong to the hot path, under the same names
    import smoothedparticles_jl_b200 as sp
    from smoothedparticles_jl_b200 import geometry as geo, io as sp_io, operators as ops
    for name in ("ParticleSystem", "ParticleField", "apply", "apply_unary", "apply_binary", "create_cell_list",
                 "assemble_vector", "wendland1", "Dwendland1", "rDwendland1", "wendland2", "Dwendland2", "rDwendland2",
                 "wendland3", "Dwendland3", "rDwendland3", "DDwendland3", "spline24", "Dspline24", "rDspline24",
                 "spline23", "Dspline23", "rDspline23"):
        assert callable(getattr(sp, name)), name
    for name in ("Squaregrid", "Hexagrid", "CubicGrid", "FacecenteredGrid", "BodycenteredGrid", "DiamondGrid", "Rectangle",
                 "Circle", "Ellipse", "Ball", "Box", "BooleanUnion", "BooleanIntersection", "BooleanDifference",
                 "Specification", "BoundaryLayer", "Transform", "Polygon", "ClosedSpline", "Ellipsoid", "Cone",
                 "RevolutionBody"):
        assert hasattr(geo, name), name
    for name in ("save_frame", "new_pvd_file", "save_pvd_file", "import_particles"):
        assert callable(getattr(sp_io, name)), name
    with pytest.raises(TypeError):
        sp.apply_unary(None, ops.balance_of_mass("wendland2", 1.0, 1.0))
    with pytest.raises(TypeError):
        sp.apply_binary(None, ops.move(0.1))


def test_every_entry_point_is_documented_and_bound_for_julia():
    # the drop-in boundary must not drift: INTEGRATION.md names every exported function, and the Julia ccall shim binds
    # every one except the instrumentation used only by bench.py / the tests
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    syms = sp.abi.declared_symbols()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    assert [s for s in syms if s not in doc] == []
    shim = open(os.path.join(root, "julia", "SmoothedParticlesB200.jl")).read()
    instrumentation = {"sp_find_field", "sp_get_sweep_neighbour_lists", "sp_last_call_ms", "sp_launch_count",
                       "sp_neighbour_list_capacity", "sp_timer_start", "sp_timer_stop"}
    assert {s for s in syms if ":" + s not in shim} <= instrumentation
    # and the Python binding table covers the header exactly
    assert sorted(sp.abi.SIGNATURES) == syms


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """include/sp_b200.h is the drop-in boundary: it must compile as C99 (pedantic) and as C++, and a C program must link
    against libsp_b200.so and see the documented error behaviour (tests/c_abi_consumer.c; with a GPU it also runs one tiny
    system through create / resize / upload / create_cell_list / destroy)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include")
    src = os.path.join(root, "tests", "c_abi_consumer.c")
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, src], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", "c++", "-I", inc, src], check=True)
    libdir = os.path.dirname(sp.abi.library_path()) if hasattr(sp.abi, "library_path") else os.path.join(root, "smoothedparticles.jl_b200")
    exe = str(tmp_path / "consumer")
    subprocess.run(["gcc", "-std=c99", "-I", inc, src, "-o", exe, "-L", libdir, "-lsp_b200", f"-Wl,-rpath,{libdir}"], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)
    assert "c-abi consumer ok" in p.stdout
