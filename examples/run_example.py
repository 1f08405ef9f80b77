"""Runs one of the reference's example scripts on the B200 engine, the way the Julia script's main() does: initial
state, prologue, time loop, a .vtp frame every `--frame-every` steps and the .pvd collection at the end (ParaView opens
`<out>/result.pvd`), plus the script's own diagnostics where a device reduction exists.

    python examples/run_example.py collapse_dry --steps 2000 --frame-every 200 --out results/collapse_dry
    python examples/run_example.py collapse3d --dr 2.5e-3 --steps 500
    python examples/run_example.py cylinder --init /path/to/examples/init/cylinder.vtp

Needs a CUDA device (there is no CPU fallback).  The configs live in smoothedparticles.jl_b200/configs.py, each a
call-for-call restatement of its script over registered operators."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smoothedparticles_jl_b200 as sp  # noqa: E402
from smoothedparticles_jl_b200 import ParticleSystem, configs, io as spio  # noqa: E402

K = sp.K
FRAME_FIELDS = {  # the save_frame! argument lists of the scripts
    "collapse_dry": ("v", "P", "rho", "type"), "collapse3d": ("v", "P", "rho", "type"), "cavity_flow": ("P", "v", "type"),
    "collapse_dry_implicit": ("v", "P", "type"), "static_container": ("v", "rho", "type"), "drop": ("v", "P", "n"),
    "collapse_symplectic": ("v", "a", "P", "rho", "rho0"), "kepler_vortex": ("v", "a", "P", "rho", "rho0"),
    "cylinder": ("v", "P", "rho", "type"), "rod": ("v", "A", "e"), "collision_2d": ("v", "P", "rho"),
}


def make_case(args):
    if args.name == "cylinder":
        if args.init:
            return configs.cylinder(args.init)
        d = np.load(os.path.join(ROOT, "tests", "golden", "cylinder_init.npz"))
        return configs.cylinder({"x": np.column_stack([d["xy"], np.zeros(len(d["xy"]))]), "type": d["type"].astype(np.float64)})
    maker = getattr(configs, args.name)
    return maker(dr=args.dr) if args.dr else maker()


def diagnostics(name, s, c):
    if name in ("collapse_dry", "collapse3d"):
        E = s.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), (c["m"], c["c"], c["rho0"], *c["g"]))[0]
        return f"E = {E:.6e}"
    if name == "rod":
        return f"E = {configs.rod_energy(s, c):.6e}"
    if name == "cylinder":
        C = configs.cylinder_force_coefficients(s, c)
        return f"C_drag = {C[0]:.4f}  C_lift = {C[1]:.4f}"
    return ""


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("name", choices=sorted(FRAME_FIELDS))
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--frame-every", type=int, default=0, help="0 = no output frames")
    ap.add_argument("--out", default=None)
    ap.add_argument("--dr", type=float, default=None, help="particle spacing (configs that take one)")
    ap.add_argument("--init", default=None, help="cylinder: path of examples/init/cylinder.vtp")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--graph", action="store_true",
                    help="record two time steps once and replay them as one CUDA graph launch per pair of steps between the "
                         "frames (sp_graph_*): a 10 k-particle 2-D step drops from ~0.2 ms to ~0.08 ms.  Not for the configs "
                         "whose step reads data back (collapse_dry_implicit: CG; cylinder: inflow respawn)")
    args = ap.parse_args()
    case = make_case(args)
    s = case.make(ParticleSystem, device=args.device)
    case.prologue(s)
    out = spio.new_pvd_file(args.out or os.path.join("results", args.name)) if args.frame_every else None
    print(f"{case.name}: {len(s)} particles, h = {case.h:g}, dt = {case.consts.get('dt', float('nan')):g}")
    t0 = time.perf_counter()
    graph = None
    k = 0
    while k <= args.steps:
        if out is not None and k % args.frame_every == 0:
            spio.save_frame(out, s, *FRAME_FIELDS[args.name])
            print(f"step {k}  N = {len(s)}  {diagnostics(args.name, s, case.consts)}", flush=True)
        # steps until the next frame (or the end)
        nxt = min(args.steps + 1, (k // args.frame_every + 1) * args.frame_every) if out is not None else args.steps + 1
        todo = nxt - k
        if args.graph and todo >= 4:
            if graph is None:
                case.step(s)                      # the ordinary way once: lazy allocations
                k, todo = k + 1, todo - 1
            pairs = todo // 2
            if pairs:
                try:
                    if graph is None:
                        graph = s.record(lambda: case.step(s), repeat=2)   # executes two steps and keeps them
                        pairs -= 1
                        k += 2
                    graph.replay(pairs)
                except sp.SpError:
                    # the system changed under the graph (e.g. the particle bound after a frame's download): record again
                    graph = s.record(lambda: case.step(s), repeat=2)
                    k += 2
                    pairs -= 1
                    graph.replay(pairs)
                k += 2 * pairs
            continue
        case.step(s)
        k += 1
    s.synchronize()
    wall = time.perf_counter() - t0
    print(f"{args.steps + 1} steps in {wall:.2f} s = {len(s) * (args.steps + 1) / wall / 1e6:.1f} M particle-updates/s (with output)")
    if out is not None:
        print("wrote", spio.save_pvd_file(out))


if __name__ == "__main__":
    main()
