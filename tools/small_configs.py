"""Step rates of the shipped example configs as shipped (2-D, 2.7-23 k particles: launch-latency bound on a GPU) and
of tests/test_collision_2d: the device path through the C ABI call by call (and through sp_run_program where a step
program exists) next to the OpenMP oracle on the host cores.  Prints one line per config and writes
gpurun_out/small_configs.json.  `--no-oracle` skips the CPU leg."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import smoothedparticles_jl_b200 as sp  # noqa: E402
from smoothedparticles_jl_b200 import ParticleSystem, configs  # noqa: E402


def cylinder_case():
    d = np.load(os.path.join(ROOT, "tests", "golden", "cylinder_init.npz"))
    xy = d["xy"]
    return configs.cylinder({"x": np.column_stack([xy, np.zeros(len(xy))]), "type": d["type"].astype(np.float64)})


MAKERS = (configs.collapse_dry, configs.cavity_flow, configs.collapse_dry_implicit, configs.collision_2d,
          configs.static_container, configs.collapse_symplectic, configs.kepler_vortex, cylinder_case, configs.rod)


def rate(case, system_cls, nsteps, warm):
    s = case.make(system_cls)
    case.prologue(s)
    for _ in range(warm):
        case.step(s)
    sync = getattr(s, "synchronize", lambda: None)
    sync()
    l0 = getattr(s, "launch_count", 0)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        case.step(s)
    sync()
    dt = time.perf_counter() - t0
    return s, dt / nsteps, (getattr(s, "launch_count", 0) - l0) / nsteps


def main():
    with_oracle = "--no-oracle" not in sys.argv
    if with_oracle:
        from oracle import oracle
        from oracle.oracle import OracleSystem
    out = []
    for maker in MAKERS:
        case = maker()
        isph = case.name == "collapse_dry_implicit"
        nsteps = 20 if isph else 200
        s, sec, launches = rate(case, ParticleSystem, nsteps, 5 if isph else 20)
        rec = {"config": case.name, "particles": len(s), "device_ms_per_step": 1e3 * sec,
               "device_updates_per_s": len(s) / sec, "launches_per_step": launches}
        if not isph and case.name != "cylinder":   # (ISPH: the CG reads back; cylinder: respawn reads back)
            try:
                g = s.record(lambda: case.step(s), repeat=2)
                s.synchronize()
                t0 = time.perf_counter()
                g.replay(100)
                s.synchronize()
                rec["graph_ms_per_step"] = 1e3 * (time.perf_counter() - t0) / 200
                rec["graph_updates_per_s"] = len(s) / (rec["graph_ms_per_step"] * 1e-3)
                g.close()
            except Exception as e:  # noqa: BLE001
                rec["graph_failed"] = str(e)[:200]
        if case.program:
            s.run_program(case.program, case.program_fields, case.program_params, 20)
            s.synchronize()
            t0 = time.perf_counter()
            s.run_program(case.program, case.program_fields, case.program_params, nsteps)
            s.synchronize()
            sec_p = (time.perf_counter() - t0) / nsteps
            rec["program_ms_per_step"] = 1e3 * sec_p
            rec["program_updates_per_s"] = len(s) / sec_p
        if with_oracle:
            so, sec_o, _ = rate(case, OracleSystem, 3 if isph else 20, 1 if isph else 3)
            rec["oracle_ms_per_step"] = 1e3 * sec_o
            rec["oracle_updates_per_s"] = len(so) / sec_o
            rec["oracle_threads"] = oracle.max_threads() if hasattr(oracle, "max_threads") else os.cpu_count()
            rec["speedup_per_call"] = sec_o / sec
            if "graph_ms_per_step" in rec:
                rec["speedup_graph"] = 1e3 * sec_o / rec["graph_ms_per_step"]
            if "program_ms_per_step" in rec:
                rec["speedup_program"] = 1e3 * sec_o / rec["program_ms_per_step"]
        out.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "small_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
