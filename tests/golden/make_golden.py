"""Generates the committed golden fixtures from the CPU oracle (run from the repo root):

    python tests/golden/make_golden.py

The reference is pure Julia and cannot run in this image, so these are NOT outputs of the Julia code: they are
outputs of the oracle restatement (pinned to the reference by tests/test_oracle_pins.py), frozen so that a later
change to the oracle or to the device code that alters results is caught.  Small on purpose (< 1 MB total).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import smoothedparticles_jl_b200 as sp  # noqa: E402
from smoothedparticles_jl_b200 import configs  # noqa: E402
from oracle import oracle  # noqa: E402
from oracle.oracle import OracleSystem  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
K = sp.K


def kernels():
    h = 0.42
    r = np.linspace(0.0, 1.25 * h, 101)
    out = {"h": h, "r": r}
    for name, kid in sp.abi.KERNEL_IDS.items():
        for kf, tag in ((K["SP_KFUN_W"], "w"), (K["SP_KFUN_DW"], "D"), (K["SP_KFUN_RDW"], "rD")):
            out[f"{name}_{tag}"] = oracle.kernel_eval(kid, kf, h, r)
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)


def steps(maker, name, nsteps, sub, **kw):
    case = maker(**kw)
    s = case.make(OracleSystem)
    case.prologue(s)
    for _ in range(nsteps):
        case.step(s)
    s.create_cell_list()
    idx = np.arange(0, len(s), sub)
    off, ids = s.neighbour_lists()
    # "count" repeats "n": a field of drop.jl is itself called n and overwrites that key
    out = {"n": len(s), "count": len(s), "nsteps": nsteps, "idx": idx, "keys": s.cell_keys()[idx], "nbr_count": np.diff(off)[idx],
           "nbr_checksum": np.array([int(ids[off[i]:off[i + 1]].sum()) for i in idx], dtype=np.int64)}
    for f in case.fields:
        out[f] = s.get(f)[idx]
    out["x"] = s.get("x")[idx]
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


def cylinder_case():
    """examples/cylinder.jl on the particles of the reference's own init/cylinder.vtp (cylinder_init.npz, made by
    tools/make_vtp_golden.py)."""
    d = np.load(os.path.join(HERE, "cylinder_init.npz"))
    xy = d["xy"]
    return configs.cylinder({"x": np.column_stack([xy, np.zeros(len(xy))]), "type": d["type"].astype(np.float64)})


if __name__ == "__main__":
    kernels()
    steps(configs.collapse_dry, "collapse_dry_5steps", 5, 7)
    steps(configs.collapse3d, "collapse3d_3steps", 3, 61, dr=1.0e-2)
    steps(configs.cavity_flow, "cavity_flow_5steps", 5, 7)
    steps(configs.collision_2d, "collision_2d_20steps", 20, 3)
    steps(configs.static_container, "static_container_5steps", 5, 9)
    steps(configs.drop, "drop_3steps", 3, 11, dr=1.2e-4)
    steps(configs.collapse_symplectic, "collapse_symplectic_10steps", 10, 5, dr=4.0e-2)
    steps(cylinder_case, "cylinder_5steps", 5, 23)
    steps(configs.rod, "rod_5steps", 5, 5)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
