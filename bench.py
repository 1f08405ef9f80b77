#!/usr/bin/env python
"""bench.py — particle-updates/s of the WCSPH step (FP64) on B200, the metric of BASELINE.json.

  python bench.py --gpus N --steps K --warmup W              our CUDA path
  python bench.py --impl reference --gpus N --steps K ...    the reference's CPU algorithm (OpenMP oracle port;
                                                             the Julia reference itself cannot run in this image)

N = 1   workload = examples/collapse3d.jl scaled to 10 M particles (BASELINE.json configs[1]): one step is one
        pass of its time loop (move, create_cell_list, balance_of_mass, find_pressure, internal_force,
        accelerate, accelerate).  Particle state (1.04 GB) is far larger than the 126 MB L2, so no flush is
        needed between steps.
N > 1   weak scaling over slabs (one process per GPU, NCCL halos/migration inside the library):
        --workload dambreak (default): the same dam break with the box depth x N along z, i.e. one copy of the
        N = 1 workload per GPU, so the per-N values are directly comparable;
        --workload box: the synthetic periodic 3-D lattice box, 25 M particles per GPU (BASELINE.json configs[4]).

One JSON line is printed by rank 0; see DESIGN.md §Measurement for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench_multi import workload_name  # noqa: E402

METRIC = "particle-updates/s (WCSPH step, FP64)"
UNIT = "particle-updates/s"
DR_10M = 9.04e-4           # examples/collapse3d.jl geometry at this dr -> 10 010 230 particles
BYTES_PER_UPDATE_STEP = 736  # algorithmic HBM bytes per particle-update of the 3-D step
# algorithmic bytes per particle of each kernel class (own fields read + written, 8 B each)
ALG_BYTES = {"internal_force": 120, "balance_of_mass": 72, "cell_list": 240, "move": 104, "find_pressure": 40,
             "accelerate": 80, "neighbour_lists": 24}
KERNEL_OF = {"internal_force": "k_sweep_list<OpInternalForceCached>",
             "balance_of_mass": "k_nbr_build_sweep<OpBalanceOfMassAux> (neighbour-list build fused with the mass sweep)"}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        if os.environ.get("SP_BENCH_NO_SAMPLER"):   # diagnostic: does the nvidia-smi poll perturb the timed region?
            return
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


def fp64_peak():
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return None


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------ N = 1
def run_single(args):
    import torch

    import smoothedparticles_jl_b200 as sp
    from smoothedparticles_jl_b200 import ParticleSystem, configs

    K = sp.K
    rank, local_rank, world = dist_env()
    dev_index = local_rank
    torch.cuda.set_device(dev_index)
    t0 = time.time()
    case = configs.collapse3d(args.dr)
    n = case.n
    gen_s = time.time() - t0

    # ---- device-resident run: `value`
    sysd = case.make(ParticleSystem, device=dev_index)
    sysd.synchronize()
    launches0 = sysd.launch_count
    sysd.run_program(case.program, case.program_fields, case.program_params, args.warmup)
    sysd.synchronize()
    sampler = ClockSampler(dev_index)
    sampler.start()
    launches1 = sysd.launch_count
    torch.cuda.synchronize()
    sysd.timer_start()
    sysd.run_program(case.program, case.program_fields, case.program_params, args.steps)
    ms = sysd.timer_stop()
    torch.cuda.synchronize()
    gpu_launches = sysd.launch_count - launches1
    n_after = len(sysd)
    ms_per_step = ms / args.steps
    value = n * args.steps / (ms * 1e-3)

    # ---- per-kernel breakdown over the same number of steps, per-call path (what a Julia host would drive)
    c = case.consts
    ops = sp.operators
    o_bom = ops.balance_of_mass("wendland3", c["m"], c["h"], c["nu"])
    o_fp = ops.find_pressure(c["dt"], c["c"], c["rho0"])
    o_if = ops.internal_force("wendland3", c["m"], c["h"], c["mu"], c["rho0"])
    o_mv = ops.move(c["dt"])
    o_ac = ops.accelerate(0.5 * c["dt"], c["g"])
    acc = {k: 0.0 for k in ("move", "cell_list", "balance_of_mass", "find_pressure", "internal_force", "accelerate")}
    # on a FRESH system: with the script's constants the 10 M scaling is only stable for ~45 steps (profiles/r2_drift.md:
    # the explicit density-diffusion term has lambda*dt = 3.1 at h = 1.8e-3, on the CPU oracle as on the device), so no
    # system is stepped for more than warm-up + K steps
    sysd.close()
    sysd = case.make(ParticleSystem, device=dev_index)
    sysd.run_program(case.program, case.program_fields, case.program_params, 3)
    sysd.synchronize()

    def timed(name, fn):
        fn()
        acc[name] += sysd.last_call_ms()

    for _ in range(args.steps):
        timed("move", lambda: sysd.apply(o_mv))
        timed("cell_list", sysd.create_cell_list)
        timed("balance_of_mass", lambda: sysd.apply(o_bom))
        timed("find_pressure", lambda: sysd.apply(o_fp))
        timed("internal_force", lambda: sysd.apply(o_if))
        timed("accelerate", lambda: sysd.apply(o_ac))
        timed("accelerate", lambda: sysd.apply(o_ac))
    clocks = sampler.stop()
    breakdown = {k: v / args.steps for k, v in acc.items()}
    dominant = max(("internal_force", "balance_of_mass"), key=lambda k: breakdown[k])
    dom_ms = breakdown[dominant]
    hbm_peak, peak_kind = _peaks()
    achieved = ALG_BYTES[dominant] * n / (dom_ms * 1e-3) / 1e9
    traffic, l1 = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(dominant)
            l1 = (tj.get("l1") or {}).get(dominant)     # ncu: what actually bounds the kernel (committed capture, not live)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": KERNEL_OF[dominant], "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_kind": peak_kind,
                "alg_bytes_per_particle": ALG_BYTES[dominant], "launch_ms": dom_ms,
                "note": "pair sweeps are bound by the L1 line-touch rate (FP64 gathers of the neighbours: one distinct 128-byte "
                        "line per clock and SM), not by HBM: see l1_ncu, fp64, step_hbm_frac and profiles/"}
    roofline["step_hbm_frac"] = BYTES_PER_UPDATE_STEP * value / (hbm_peak * 1e9)
    if l1:
        roofline["l1_ncu"] = l1
    pk = fp64_peak()
    if pk:
        roofline["fp64"] = {"peak_dfma_per_s": pk["dfma_per_s"], "peak_dadd_per_s": pk["dadd_per_s"],
                            "peak_tflops_fma": pk["fp64_tflops_fma"]}
    sysd.close()
    del sysd

    # ---- the same initial state built by the device generator (sp_generate_particles) instead of numpy + upload
    device_setup_s = None
    try:
        tmp = case.make_on_device(ParticleSystem, device=dev_index)   # warm-up (pool growth, module load)
        tmp.close()
        t0 = time.perf_counter()
        tmp = case.make_on_device(ParticleSystem, device=dev_index)
        tmp.synchronize()
        device_setup_s = time.perf_counter() - t0
        assert len(tmp) == n
        tmp.close()
        del tmp
    except Exception as e:  # noqa: BLE001 - the generator is not on the measured path
        device_setup_s = f"failed: {e}"

    # ---- end to end through the C ABI with HOST buffers: upload from pinned memory, K steps driven call by
    # call with a per-step device->host diagnostic (total energy), download of the result fields.
    e2e = run_e2e(case, args, dev_index)

    # ---- BASELINE configs[4] on this one GPU: the periodic box, 292^3 particles, as one slab exchanging ghosts with
    # itself — the N = 1 point of its weak-scaling curve (the N > 1 lines carry the same key)
    box = None
    if not args.no_box:
        try:
            from bench_multi import box_line
            box = box_line(args, UNIT, ClockSampler)
        except Exception as e:  # noqa: BLE001 - the extra line must not take the headline down
            box = {"failed": str(e)[:300]}

    # ---- CPU baseline on a bounded sample (rank 0 only)
    cpu = cpu_baseline(case, sample_budget_s=args.cpu_budget) if not args.no_cpu else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name("dambreak", 1, args.dr, args.per_gpu),
                   "particles": n, "particles_after": n_after, "h": case.h, "cells": int(np.prod(_key_lim(case))),
                   "l2": "state 1.04 GB >> 126 MB L2, no flush needed", "driver": "sp_run_program (fused step loop)",
                   "stability": "the script's constants are stable for ~45 steps at this resolution (CPU oracle and device "
                                "alike, profiles/r2_drift.md): every timed system runs warm-up + K steps from the initial state",
                   "setup_s": round(gen_s, 1),
                   "device_setup_s": (round(device_setup_s, 4) if isinstance(device_setup_s, float) else device_setup_s)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches), "roofline": roofline,
        "breakdown_ms": breakdown, "cpu_baseline": cpu, "box": box,
    }
    emit(line)


def _key_lim(case):
    lo, hi, h = case.domain.lo, case.domain.hi, case.h
    return [int(np.floor(hi[a] / h) - np.floor(lo[a] / h) + 1) for a in range(3)]


def run_e2e(case, args, dev_index):
    import torch

    import smoothedparticles_jl_b200 as sp
    from smoothedparticles_jl_b200 import ParticleSystem

    K = sp.K
    n = case.n
    c = case.consts
    names_in = ["x", "v", "rho", "type"]                  # non-zero initial fields (the rest start at zero)
    names_out = ["x", "v", "rho", "P"]                    # what save_frame! would fetch, plus positions
    host_in = {}
    for nm in names_in:
        nc = 3 if nm in ("x", "v") else 1
        src = case.init.get(nm)
        t = torch.empty((n, nc) if nc > 1 else (n,), dtype=torch.float64, pin_memory=True)
        if src is not None:
            t.copy_(torch.from_numpy(np.ascontiguousarray(src)))
        else:
            t.zero_()
        host_in[nm] = t
    host_out = {nm: torch.empty((n, 3) if nm in ("x", "v") else (n,), dtype=torch.float64, pin_memory=True)
                for nm in names_out}
    ops = sp.operators
    o_bom = ops.balance_of_mass("wendland3", c["m"], c["h"], c["nu"])
    o_fp = ops.find_pressure(c["dt"], c["c"], c["rho0"])
    o_if = ops.internal_force("wendland3", c["m"], c["h"], c["mu"], c["rho0"])
    o_mv = ops.move(c["dt"])
    o_ac = ops.accelerate(0.5 * c["dt"], c["g"])
    pe = (c["m"], c["c"], c["rho0"], *c["g"])
    import ctypes as C

    phases = {}

    def job():
        tp = [time.perf_counter()]

        def mark(name):
            tp.append(time.perf_counter())
            phases[name] = tp[-1] - tp[-2]

        s = ParticleSystem(case.fields, case.domain, case.h, device=dev_index)
        s.resize(n)
        s.synchronize()
        mark("create_s")
        for nm, t in host_in.items():
            s.upload_raw(nm, C.cast(t.data_ptr(), C.POINTER(C.c_double)), n, K["SP_LAYOUT_AOS"])
        mark("upload_s")
        energy = 0.0
        energies = []
        for _ in range(args.steps):
            s.apply(o_mv)
            s.create_cell_list()
            s.apply(o_bom)
            s.apply(o_fp)
            s.apply(o_if)
            s.apply(o_ac)
            s.apply(o_ac)
            energy = s.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), pe)[0]   # D2H every step
            energies.append(energy)
        mark("steps_s")
        for nm, t in host_out.items():
            s.download_raw(nm, C.cast(t.data_ptr(), C.POINTER(C.c_double)), len(s), K["SP_LAYOUT_AOS"])
        s.synchronize()
        mark("download_s")
        s.close()
        mark("destroy_s")
        phases["energy_first"] = energies[0] if energies else 0.0
        return energy

    job()  # warm-up (allocations, page-ins, lazy module load)
    best, best_phases, energy = None, None, 0.0
    for _ in range(3):  # best of 3 jobs: the job is ~0.1 s and host-side noise (page faults, other tenants) is visible
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        energy = job()
        torch.cuda.synchronize()
        dt_job = time.perf_counter() - t0
        if best is None or dt_job < best:
            best, best_phases = dt_job, dict(phases)
    dt = best
    phases.update(best_phases)
    h2d = sum(t.numel() * 8 for t in host_in.values())
    d2h = sum(t.numel() * 8 for t in host_out.values())
    return {"value": n * args.steps / dt, "unit": UNIT, "steps": args.steps, "h2d_bytes_per_step": h2d / args.steps,
            "d2h_bytes_per_step": d2h / args.steps + 24, "seconds": dt, "phases": {k: round(v, 4) for k, v in phases.items() if k.endswith("_s")},
            "what": f"one job = sp_create + upload of x,v,rho,type from pinned host memory, K = {args.steps} steps driven "
                    "call by call through the C ABI with a per-step energy read-back, download of x,v,rho,P, sp_destroy; "
                    "best of 3 jobs; the bulk upload/download is amortised over K (particles stay resident in HBM), so "
                    "this value moves with K",
            "energy": energy,
            "energy_drift": ((energy - phases["energy_first"]) / abs(phases["energy_first"])
                             if phases.get("energy_first") else None)}


def host_cores() -> int:
    """All the cores this process may run on (the launcher's OMP_NUM_THREADS=1 under torch.distributed.run is ignored)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def time_oracle(case, warmup, steps):
    """ONE protocol for both CPU legs: the OpenMP restatement of the reference's algorithm (oracle/) on all host cores,
    `warmup` untimed steps (first touch of the pages, thread pool start), then `steps` timed steps of the step program."""
    from oracle import oracle
    from oracle.oracle import OracleSystem

    threads = host_cores()
    oracle.load().so_set_threads(threads)
    s = case.make(OracleSystem)
    n = len(s)
    if warmup:
        s.run_program(case.program, case.program_fields, case.program_params, warmup)
    t = s.run_program(case.program, case.program_fields, case.program_params, steps)
    s.close()
    return n, t, threads


def cpu_baseline(case, sample_budget_s=20.0):
    """The reference's algorithm on the host cores, same workload, bounded sample: 2 warm-up steps + as many timed steps
    as fit the budget (at most 10)."""
    from oracle import oracle
    from oracle.oracle import OracleSystem

    threads = host_cores()
    oracle.load().so_set_threads(threads)
    s = case.make(OracleSystem)
    n = len(s)
    t_warm = s.run_program(case.program, case.program_fields, case.program_params, 2)
    per = t_warm / 2
    steps = int(max(2, min(10, sample_budget_s / max(per, 1e-3))))
    t = s.run_program(case.program, case.program_fields, case.program_params, steps)
    s.close()
    return {"value": n * steps / t, "unit": UNIT, "cores": threads, "cpu": cpu_model(), "kind": "port",
            "sample": f"{steps} step(s) of the full {n}-particle workload after 2 warm-up steps, "
                      f"OpenMP oracle (C++ restatement of the reference; Julia is not installed)",
            "ms_per_step": 1e3 * t / steps}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm on ALL host cores (OpenMP port of it; the Julia original cannot
    run here).  N > 1: rank 0 alone runs; the weak-scaled workload is N copies of one GPU's share, and the CPU's rate in
    particle-updates/s does not depend on how many copies it works through (linear-time algorithm), so one share is
    timed — the bounded sample — and its rate is the whole-job rate of the CPU."""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    from smoothedparticles_jl_b200 import configs

    if args.gpus > 1 and args.workload == "box":
        case = configs.lattice_box(128, jitter=0.1)
        share = "a 128^3 block of the same lattice"
    else:
        case = configs.collapse3d(args.dr)
        share = "one GPU's share (the 10 M-particle dam break)" if args.gpus > 1 else "the full workload"
    n, t, threads = time_oracle(case, args.warmup, args.steps)
    value = n * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload if args.gpus > 1 else "dambreak", args.gpus, args.dr, args.per_gpu),
                   "particles": n * (args.gpus if args.gpus > 1 and args.workload != "box" else 1),
                   "particles_timed": n,
                   "note": "each step is one full time step of the timed particles on the host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "cpu": cpu_model(), "kind": "port",
                         "sample": f"{args.steps} steps of {share} ({n} particles) after {args.warmup} warm-up steps, "
                                   f"OpenMP oracle port on {threads} threads (the Julia reference cannot run here: no "
                                   f"Julia runtime); particle-updates/s of the CPU is independent of the number of copies"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ N > 1
def run_multi(args):
    from bench_multi import run_multi as _run
    _run(args, METRIC, UNIT, ClockSampler, _peaks, emit)


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries write to fd 1 on their own (NCCL prints its version banner there when NCCL_DEBUG asks for it): everything
    # but the JSON line goes to stderr, so stdout carries exactly one line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    wd = os.environ.get("SP_BENCH_WATCHDOG")   # debugging aid: dump every thread's Python stack and exit after N seconds
    if wd:
        import faulthandler
        faulthandler.dump_traceback_later(float(wd), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20,
                    help="timed steps (with the script's constants the 10 M scaling is stable for ~45 steps: warm-up + steps "
                         "should stay below that, see profiles/r2_drift.md)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dr", type=float, default=DR_10M)
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--workload", default="dambreak", choices=["dambreak", "box"],
                    help="N > 1: 'dambreak' = collapse3d with the box depth x N (one N=1 workload per GPU); "
                         "'box' = periodic lattice box, --per-gpu^3 particles per GPU (BASELINE configs[4])")
    ap.add_argument("--per-gpu", type=int, default=292, help="lattice side per GPU for --workload box (292^3 = 24.9 M)")
    ap.add_argument("--no-box", action="store_true", help="skip the extra `box` line (BASELINE configs[4] at this N)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not os.environ.get("SP_BENCH_ALLOW_SHORT"):
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus == 1:
        run_single(args)
    else:
        try:
            run_multi(args)
        except BaseException:
            # a rank that fails must not sit in a destructor waiting for a collective its peers will never join
            import traceback
            traceback.print_exc()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(1)


if __name__ == "__main__":
    main()
