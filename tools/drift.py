"""Trajectory and energy drift over many steps (north_star: "bounded trajectory and energy drift over N steps").

  A. examples/collapse3d.jl at dr = 2.2e-3 (~0.8 M particles): device (step program, CUDA graph replay) against the CPU oracle
     running the same loop, compared every 50 steps for 200 steps: max|a-b|/max|b| of x, v, rho, and the particle count.
  B. the same script at dr = 1.5e-3 (2.3 M particles) for 2000 steps on the device: total energy every 200 steps, count.
     (NOT the 10 M-particle bench workload: with the script's constants the explicit density-diffusion term
     2*nu*(rho_p - rho_q) is unstable for h < 2.8e-3 — lambda*dt = 4e-4*14/h > 2 with the quintic Wendland kernel — so the
     10 M case (h = 1.8e-3) blows up after ~45 steps, on the CPU oracle exactly as on the device; see profiles/r2_drift.md.)
Writes gpurun_out/drift.json."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smoothedparticles_jl_b200 as sp  # noqa: E402
from smoothedparticles_jl_b200 import ParticleSystem, configs  # noqa: E402
from oracle.oracle import OracleSystem  # noqa: E402

K = sp.K
out = {}
case = configs.collapse3d(2.2e-3)
c = case.consts
pe = (c["m"], c["c"], c["rho0"], *c["g"])
dev, ora = case.make(ParticleSystem), case.make(OracleSystem)
rows = []
for block in range(4):
    dev.run_program(case.program, case.program_fields, case.program_params, 50)
    ora.run_program(case.program, case.program_fields, case.program_params, 50)
    row = {"steps": 50 * (block + 1), "n_dev": len(dev), "n_ora": len(ora)}
    for nm in ("x", "v", "rho"):
        a, b = dev.get(nm), ora.get(nm)
        row[nm] = float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    Ed = dev.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), pe)[0]
    Eo = ora.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), pe)[0]
    row["energy_rel_diff"] = abs(Ed - Eo) / abs(Eo)
    rows.append(row)
    print(row, flush=True)
out["A_device_vs_oracle"] = {"particles": case.n, "dr": 2.2e-3, "rows": rows}
dev.close()
ora.close()

case = configs.collapse3d(1.5e-3)
c = case.consts
pe = (c["m"], c["c"], c["rho0"], *c["g"])
dev = case.make(ParticleSystem)
E0 = dev.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), pe)[0]
rows = [{"steps": 0, "E": E0, "n": len(dev)}]
t0 = time.perf_counter()
for block in range(10):
    dev.run_program(case.program, case.program_fields, case.program_params, 200)
    E = dev.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), pe)[0]
    rows.append({"steps": 200 * (block + 1), "E": E, "rel_drift": (E - E0) / abs(E0), "n": len(dev)})
    print(rows[-1], flush=True)
out["B_2M_2000_steps"] = {"particles": case.n, "rows": rows, "wall_s": time.perf_counter() - t0,
                           "vmax": float(np.max(np.abs(dev.get("v"))))}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/drift.json", "w"), indent=1)
