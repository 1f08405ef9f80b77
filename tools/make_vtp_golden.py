"""Golden summary of the one VTK file the reference ships (examples/init/cylinder.vtp, written by the reference's own
save_frame! through WriteVTK.jl): read it with OUR reader and record counts, names and checksums.  Run in the build
container (needs /root/reference); the summary is committed as tests/golden/cylinder_vtp_summary.json and pins
smoothedparticles.jl_b200/io.py::read_vtp against a file our writer did not produce."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("sp_io", os.path.join(ROOT, "smoothedparticles.jl_b200", "io.py"))
io = importlib.util.module_from_spec(spec)
spec.loader.exec_module(io)

SRC = "/root/reference/examples/init/cylinder.vtp"
pts, fields = io.read_vtp(SRC)
summary = {"source": "examples/init/cylinder.vtp", "n": int(pts.shape[0]),
           "points_sha256": hashlib.sha256(np.ascontiguousarray(pts).tobytes()).hexdigest(),
           "points_min": pts.min(axis=0).tolist(), "points_max": pts.max(axis=0).tolist(),
           "fields": {k: {"shape": list(v.shape), "sha256": hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest(),
                          "min": float(v.min()), "max": float(v.max()), "sum": float(v.sum())} for k, v in fields.items()}}
out = os.path.join(ROOT, "tests", "golden", "cylinder_vtp_summary.json")
json.dump(summary, open(out, "w"), indent=1)
print(json.dumps(summary, indent=1))

# the same particles as an initial-state fixture for configs.cylinder() (examples/cylinder.jl:86-89 imports this
# file): exact coordinates (z is identically 0) and the particle types; checked against the summary above in
# tests/test_cylinder_cpu.py
assert np.all(pts[:, 2] == 0.0)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cylinder_init.npz"), xy=pts[:, :2].copy(),
                    type=fields["type"].astype(np.uint8))
