"""Step rates of the shipped 2-D configs (10-23 k particles: launch-latency bound) and of tests/test_collision_2d."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smoothedparticles_jl_b200 as sp  # noqa: E402
from smoothedparticles_jl_b200 import ParticleSystem, configs  # noqa: E402

for maker in (configs.collapse_dry, configs.cavity_flow, configs.collision_2d, configs.static_container,
              configs.collapse_dry_implicit):
    case = maker()
    s = case.make(ParticleSystem)
    case.prologue(s)
    for _ in range(20):
        case.step(s)
    s.synchronize()
    l0 = s.launch_count
    nsteps = 200 if case.name != "collapse_dry_implicit" else 20
    t0 = time.perf_counter()
    for _ in range(nsteps):
        case.step(s)
    s.synchronize()
    dt = time.perf_counter() - t0
    line = f"{case.name:24s} n={len(s):6d}  per-call: {1e3 * dt / nsteps:7.3f} ms/step  {len(s) * nsteps / dt / 1e6:7.2f} M updates/s  launches/step={(s.launch_count - l0) / nsteps:.0f}"
    if case.program:
        s.run_program(case.program, case.program_fields, case.program_params, 20)
        s.synchronize()
        t0 = time.perf_counter()
        s.run_program(case.program, case.program_fields, case.program_params, nsteps)
        s.synchronize()
        dt = time.perf_counter() - t0
        line += f" | run_program: {1e3 * dt / nsteps:7.3f} ms/step {len(s) * nsteps / dt / 1e6:7.2f} M updates/s"
    print(line, flush=True)
