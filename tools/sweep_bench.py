"""Times the pair sweeps (and the rest of the 3-D WCSPH step) call by call on the 10 M dam break or a lattice
block; used under ncu for the per-kernel captures in profiles/."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smoothedparticles_jl_b200 as sp  # noqa: E402
from smoothedparticles_jl_b200 import ParticleSystem, configs, operators as ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="dambreak")
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--dr", type=float, default=9.04e-4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warm", type=int, default=5)
ap.add_argument("--unfused", action="store_true", help="list build and first replay as two kernels")
args = ap.parse_args()
case = configs.collapse3d(args.dr) if args.case == "dambreak" else configs.lattice_box(args.n, jitter=0.1)
c = case.consts
s = case.make(ParticleSystem)
o_bom = ops.balance_of_mass("wendland3", c["m"], c["h"], c["nu"])
o_fp = ops.find_pressure(c["dt"], c["c"], c["rho0"])
o_if = ops.internal_force("wendland3", c["m"], c["h"], c["mu"], c["rho0"])
o_mv = ops.move(c["dt"])
o_ac = ops.accelerate(0.5 * c["dt"], c["g"])
s.run_program(case.program, case.program_fields, case.program_params, args.warm)   # develop some disorder
acc = {}


def timed(name, fn):
    fn()
    acc.setdefault(name, []).append(s.last_call_ms())


for _ in range(args.steps):
    timed("move", lambda: s.apply(o_mv))
    timed("cell_list", s.create_cell_list)
    timed("balance_of_mass", lambda: s.apply(o_bom, unfused_build=args.unfused))
    timed("find_pressure", lambda: s.apply(o_fp))
    timed("internal_force", lambda: s.apply(o_if))
    timed("accelerate", lambda: s.apply(o_ac))
    timed("accelerate2", lambda: s.apply(o_ac))
n = len(s)
tot = sum(np.mean(v) for v in acc.values())
print(f"case={case.name} n={n} step={tot:.3f} ms  -> {n / tot / 1e3:.1f} M updates/s")
for k, v in acc.items():
    print(f"  {k:16s} {np.mean(v):8.3f} ms   {n / np.mean(v) / 1e6:8.2f} G particles/s")
off, ids = (None, None)
if os.environ.get("SP_COUNT_NBRS"):
    off, ids = s.neighbour_lists()
    print("  mean neighbours", np.mean(np.diff(off)), "max", np.max(np.diff(off)))
