"""Physical validation of the 2-D dam break against the data the reference ships for exactly this comparison:
examples/collapse_dry.jl:232-249 (make_plot) overlays the computed wave front X(t) and column height H(t) on
examples/reference/dambreak_{X,H}_{Violeau,Koshizuka}.csv (Violeau's SPH result, "Fluid Mechanics and the SPH Method"
p. 484, and the experiment of Koshizuka & Oka 1996).  Those four tables (72 points: data, not code) are committed as
tests/golden/dambreak_tables.json (`--tables` regenerates the file from the reference checkout) so the script also runs
on the GPU box, where /root/reference does not exist.

Runs the script's time loop (configs.collapse_dry: the statements of collapse_dry.jl:203-211, dr = 1.5e-2 as shipped)
to the dimensionless time t*sqrt(2 g) = 3.0 the reference plots, with get_globals (:177-191) evaluated every 50 steps,
and reports the deviation from the tables.  This is NOT bit parity — no Julia runtime exists here, so the oracle stays
unpinned in that sense — it is the external anchor the reference itself uses: a restatement with a wrong kernel
normalisation, pressure law, wall treatment or time scheme does not land on these curves.

    python tools/dambreak_validation.py oracle            # CPU oracle (here or on the GPU box)
    python tools/dambreak_validation.py device            # CUDA path through the C ABI (GPU box)
    python tools/dambreak_validation.py oracle device     # both, and the difference between them
    python tools/dambreak_validation.py --isph oracle     # the ISPH dam break (collapse_dry_implicit.jl) against the same tables
Writes gpurun_out/dambreak_validation_<backends>.json."""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import smoothedparticles_jl_b200 as sp  # noqa: E402,F401
from smoothedparticles_jl_b200 import configs  # noqa: E402


def load_tables():
    """The four tables (time = t*sqrt(2 g), X = front / column width, H = height / column height) from the committed copy
    tests/golden/dambreak_tables.json; `--tables` rewrites that file from the reference checkout (this container only)."""
    path = os.path.join(ROOT, "tests", "golden", "dambreak_tables.json")
    ref = "/root/reference/examples/reference"
    if "--tables" in sys.argv:
        out = {}
        for q in ("X", "H"):
            for who in ("Violeau", "Koshizuka"):
                rows = open(os.path.join(ref, f"dambreak_{q}_{who}.csv")).read().strip().splitlines()[1:]
                pts = sorted(tuple(float(v) for v in r.split(",")) for r in rows if r.strip())
                out[f"{q}_{who}"] = pts
        json.dump(out, open(path, "w"), indent=0)
        print("wrote", path)
        sys.exit(0)
    return {k: np.asarray(v) for k, v in json.load(open(path)).items()}


def globals_of(sys_, c):
    """get_globals, collapse_dry.jl:177-191 (X and H only)."""
    x = sys_.get("x")
    fluid = sys_.get("type") == 0.0
    X = float(np.max(x[fluid, 0])) / c["width"] if fluid.any() else 0.0
    col = fluid & (x[:, 0] < 2.0) & (x[:, 0] > c["h"])
    H = float(np.max(x[col, 1])) / c["height"] if col.any() else 0.0
    return X, H


def run(system_cls, every=50, t_star_end=3.0, isph=False):
    """WCSPH: collapse_dry.jl as shipped.  isph=True: collapse_dry_implicit.jl as shipped (dr = 1e-2, 23 172 particles, loop
    :205-233 with the matrix-free CG in place of assemble_matrix + cg), whose make_plot (:242-255) uses the same tables."""
    case = configs.collapse_dry_implicit() if isph else configs.collapse_dry()
    c = dict(case.consts)
    c.setdefault("width", 1.0)
    c.setdefault("height", 2.0)
    scale = math.sqrt(2.0 * abs(c["g"][1]))
    nsteps = int(math.ceil(t_star_end / scale / c["dt"])) + every
    s = case.make(system_cls)
    case.prologue(s)
    ts, Xs, Hs = [], [], []
    t0 = time.perf_counter()
    for k in range(nsteps + 1):
        if not isph:
            case.step(s)
        if k % every == 0:   # collapse_dry.jl samples after the step with index k (:212-222), the ISPH script before it (:206-217)
            X, H = globals_of(s, c)
            ts.append(k * c["dt"] * scale)
            Xs.append(X)
            Hs.append(H)
        if isph:
            case.step(s)
    wall = time.perf_counter() - t0
    return {"t": ts, "X": Xs, "H": Hs, "steps": nsteps + 1, "particles": len(s), "wall_s": wall}


def compare(curve, tables):
    out = {}
    t = np.asarray(curve["t"])
    for name, tab in tables.items():
        q = name[0]
        if not len(tab):
            out[name] = {"points": 0}
            continue
        y = np.interp(tab[:, 0], t, np.asarray(curve[q]))
        keep = tab[:, 0] <= t[-1]
        # the front stops at the far wall (X = box_width / column width = 4): compare while it is under way
        d = (y - tab[:, 1])[keep]
        if not d.size:
            out[name] = {"points": 0}
            continue
        out[name] = {"points": int(keep.sum()), "max_abs_dev": float(np.max(np.abs(d))), "mean_dev": float(np.mean(d)),
                     "rms_dev": float(np.sqrt(np.mean(d * d)))}
    return out


def main():
    tables = load_tables()
    res = {"config": "collapse_dry.jl as shipped (dr = 1.5e-2), loop :203-211, to t*sqrt(2g) = 3.0",
           "tables": "examples/reference/dambreak_{X,H}_{Violeau,Koshizuka}.csv (collapse_dry.jl:232-249)"}
    curves = {}
    isph = "--isph" in sys.argv
    if isph:
        res["config"] = "collapse_dry_implicit.jl as shipped (dr = 1e-2), loop :205-233, to t*sqrt(2g) = 3.0"
        res["tables"] = "examples/reference/dambreak_{X,H}_{Violeau,Koshizuka}.csv (collapse_dry_implicit.jl:242-255)"
    every = 10 if isph else 50
    if "oracle" in sys.argv:
        from oracle.oracle import OracleSystem
        curves["oracle"] = run(OracleSystem, every=every, isph=isph)
    if "device" in sys.argv:
        from smoothedparticles_jl_b200 import ParticleSystem
        curves["device"] = run(ParticleSystem, every=every, isph=isph)
    for who, cv in curves.items():
        res[who] = {"steps": cv["steps"], "particles": cv["particles"], "wall_s": cv["wall_s"], "vs_tables": compare(cv, tables),
                    "curve": {"t": cv["t"], "X": cv["X"], "H": cv["H"]}}
        print(who, json.dumps(res[who]["vs_tables"]), f"{cv['wall_s']:.1f} s", flush=True)
    if len(curves) == 2:
        a, b = curves["oracle"], curves["device"]
        res["device_minus_oracle"] = {"max_abs_dX": float(np.max(np.abs(np.asarray(a["X"]) - np.asarray(b["X"])))),
                                      "max_abs_dH": float(np.max(np.abs(np.asarray(a["H"]) - np.asarray(b["H"]))))}
        print("device - oracle", json.dumps(res["device_minus_oracle"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    name = "dambreak_validation_" + ("isph_" if isph else "") + "_".join(curves) + ".json"
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", name), "w"))


if __name__ == "__main__":
    main()
