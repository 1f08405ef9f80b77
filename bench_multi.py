"""bench.py --gpus N (N > 1): weak scaling of the 3-D WCSPH step on the synthetic periodic lattice box
(BASELINE.json configs[4]): `per_gpu`^3 particles per GPU (292^3 = 24.9 M), slab-decomposed along z, one
process per GPU, NCCL ghost halos / migration inside the library, gloo for the control plane."""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np


def make_rank_particles(per_gpu: int, rank: int, world: int, dr: float = 1.0, jitter: float = 0.1, seed: int = 1234):
    """This rank's share of the global nx x ny x (nz*world) lattice: planes k in [rank*nz, (rank+1)*nz)."""
    nx = ny = nz = per_gpu
    c = 50.0
    I, J, Kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(rank * nz, (rank + 1) * nz), indexing="ij")
    n = I.size
    x = np.empty((n, 3))
    x[:, 0] = I.ravel() * dr
    x[:, 1] = J.ravel() * dr
    x[:, 2] = (Kk.ravel() + 0.5) * dr
    rng = np.random.Generator(np.random.Philox(key=seed + 1000 * rank))
    x += rng.uniform(-jitter * dr, jitter * dr, size=(n, 3))
    v = rng.uniform(-0.01 * c, 0.01 * c, size=(n, 3))
    perm = np.random.Generator(np.random.Philox(key=99 + rank)).permutation(n)  # unsorted input order
    return x[perm], v[perm]


def workload_name(workload: str, world: int, dr: float, per: int) -> str:
    """The `config.workload` string: the same for our arm and for --impl reference at the same N."""
    if workload == "box":
        return (f"periodic 3-D lattice box, {per}^3 particles per GPU x {world} GPU(s), slab-decomposed along z "
                f"(h = 2 dr, jitter 0.1 dr, collapse3d step)")
    if world == 1:
        return "examples/collapse3d.jl dam break scaled to 10 M particles (dr=%g)" % dr
    return (f"examples/collapse3d.jl dam break, dr={dr:g}, box extruded along z to {world} x the particle count of the N=1 "
            f"workload, slab-decomposed along z with count-balanced cuts")


def init_control_plane():
    """gloo process group for the control plane (NCCL id broadcast, max over ranks); world 1 without torchrun works too."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    if not dist.is_initialized():
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return dist, rank, world


def run_slab_workload(args, workload, UNIT, ClockSampler, e2e=True, breakdown=True):
    """One weak-scaling workload on `world` slabs: returns a dict (every rank), rank 0's is complete."""
    import torch

    import smoothedparticles_jl_b200 as sp
    from smoothedparticles_jl_b200 import geometry as geo, operators as ops, slab

    K = sp.K
    dist, rank, world = init_control_plane()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)

    def fresh_id():
        # one NCCL unique id per communicator: rank 0 draws it, gloo carries the 128 bytes to the others
        ids = [slab.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    per = args.per_gpu
    rho0, c, mu, nu = 1000.0, 50.0, 8.4e-4, 1.0e-4
    fields = {"v": 3, "Dv": 3, "P": 1, "rho": 1, "Drho": 1, "type": 1}
    t0 = time.time()
    if workload == "box":
        # BASELINE.json configs[4]: periodic lattice box, per^3 particles per GPU
        dr = 1.0
        h = 2.0 * dr
        g = (0.0, 0.0, -9.8)
        Lz = per * world * dr
        dom = geo.Box(-h, -h, 0.0, (per - 1) * dr + h, (per - 1) * dr + h, Lz * (1 - 1e-13))
        periodic = True
        x, v = make_rank_particles(per, rank, world, dr)
        typ = np.zeros(len(x))
    else:
        # examples/collapse3d.jl at dr = args.dr (10 M particles per unit depth), box depth x N along z — the
        # direction the dam break is invariant in — so every GPU holds one copy of the N = 1 workload
        from smoothedparticles_jl_b200 import configs
        dr = args.dr
        h = 2.0 * dr
        g = (0.0, 0.0, -9.8)
        wall = 2.5 * dr
        # WEAK scaling: every GPU gets the particle count of the N = 1 workload.  The two end walls exist once per job,
        # not once per GPU, so the box is a little deeper than N x 0.15, and the slabs are cut by particle count, not by
        # layer count (the end layers are full planes of wall particles, ~4x as dense as a fluid layer).  Both follow
        # from two numbers every rank measures on a thin slice of the lattice: particles per interior cell layer (f)
        # and per end-wall layer (w).
        probe = configs.collapse3d(dr, depth_scale=1.0, z_range=(-wall - dr, 3.0 * h + 0.5 * dr))
        pl = np.floor(probe.init["x"][:, 2] / h).astype(np.int64) - int(np.floor(-wall / h))
        head = np.bincount(pl, minlength=5)[:5]          # layers 0..4 from the lower domain face
        f_layer = float(head[4])                          # an interior layer (two lattice planes of fluid + side walls)
        end_layers = head[:2].astype(float)               # the layers that hold the end wall
        del probe
        n_wall_end = float(end_layers.sum())

        def total_for(depth):
            gl = int(np.floor((depth + wall) / h)) - int(np.floor(-wall / h)) + 1
            return 2.0 * n_wall_end + f_layer * (gl - 4), gl

        n1, _ = total_for(0.15)
        depth = 0.15 * world
        if world > 1:
            # the depth at which the job holds world x n1 particles (whole cell layers)
            gl_target = 4 + int(round((world * n1 - 2.0 * n_wall_end) / f_layer))
            depth = (gl_target - 1 + int(np.floor(-wall / h))) * h - wall + 0.5 * h
        dom = geo.Box(-wall, -wall, -wall, 0.584 + wall, 0.35 + wall, depth + wall)
        gphase = int(np.floor(dom.lo[2] / h))
        glim = int(np.floor(dom.hi[2] / h)) - gphase + 1
        layer_counts = np.full(glim, f_layer)
        layer_counts[:2] = end_layers
        # the far end: how the wall falls onto the cell layers depends on the depth, so it is measured as well
        probe = configs.collapse3d(dr, depth_scale=depth / 0.15, z_range=((gphase + glim - 4) * h, depth + wall + dr))
        tl = np.floor(probe.init["x"][:, 2] / h).astype(np.int64) - gphase
        layer_counts[glim - 4:] = np.bincount(tl, minlength=glim)[glim - 4:]
        del probe
        cuts = slab.balanced_cuts(layer_counts, world) if world > 1 else None
        c0, c1 = (cuts[rank], cuts[rank + 1]) if cuts else (0, glim)
        zlo, zhi = (gphase + c0) * h - dr, (gphase + c1) * h + dr      # generous: ownership is decided below
        case = configs.collapse3d(dr, depth_scale=depth / 0.15, z_range=(zlo, zhi))
        x = case.init["x"]
        cell = np.floor(x[:, 2] / h).astype(np.int64) - gphase
        mine = (cell >= c0) & (cell < c1)
        x = x[mine]
        typ = case.init["type"][mine]
        v = np.zeros_like(x)
        periodic = False
        assert np.allclose(case.domain.lo, dom.lo) and np.allclose(case.domain.hi, dom.hi), (case.domain, dom)
        dom = case.domain
    slab_cuts = cuts if workload != "box" else None
    wname = workload_name(workload, world, args.dr, per)
    m = rho0 * dr ** 3
    dt = 0.1 * h / c
    n_local = len(x)
    gen_s = time.time() - t0
    o = dict(bom=ops.balance_of_mass("wendland3", m, h, nu), fp=ops.find_pressure(dt, c, rho0),
             force=ops.internal_force("wendland3", m, h, mu, rho0), move=ops.move(dt), acc=ops.accelerate(0.5 * dt, g))

    def make_system():
        return slab.SlabSystem(fields, dom, h, rank, world, fresh_id(), periodic=periodic, device=local, cuts=slab_cuts)

    sysd = make_system()
    assert np.all(sysd.owns(x)), "generator and slab partition disagree"
    sysd.add_particles(x=x, v=v, rho=np.full(n_local, rho0), type=typ)
    # the step loop is issued from inside the library (sp_run_program on a slab system: slab rebuild with migration
    # and ghost halos, sweeps) — the same driver as the N = 1 `value`
    prog = K["SP_PROGRAM_WCSPH_3D"]
    prog_fields = ("x", "v", "Dv", "rho", "Drho", "P", "type")
    prog_params = (float(K["SP_KERNEL_WENDLAND3"]), m, h, 2 * nu, dt, c * c, rho0, mu, *g)
    sysd.run_program(prog, prog_fields, prog_params, args.warmup)
    sysd.synchronize()
    launches0 = sysd.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    sysd.timer_start()
    sysd.run_program(prog, prog_fields, prog_params, args.steps)
    ms = sysd.timer_stop()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = sysd.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t[0])
    n_tot = int(sysd.allreduce([sysd.n_owned])[0])
    cnts = np.zeros(world)
    cnts[rank] = sysd.n_owned
    sysd_counts = sysd.allreduce(cnts)
    n_ghost = len(sysd) - sysd.n_owned
    value = n_tot * args.steps / (ms_max * 1e-3)
    out = {"workload": wname, "value": value, "ms_per_step": ms_max / args.steps, "particles": n_tot,
           "particles_per_gpu": n_local, "ghosts_per_gpu": int(n_ghost), "setup_s": round(gen_s, 1), "clocks": clocks,
           "particles_per_gpu_all": [int(v) for v in sysd_counts],
           "gpu_launches": int(gpu_launches), "n_slots": n_local + int(n_ghost)}

    if breakdown:
        # per-kernel breakdown on this rank (same steps, per-call timing) — on a fresh system: with the script's constants
        # the dam break at this resolution is only stable for ~45 steps (profiles/r2_drift.md)
        sysd.close()
        sysd = make_system()
        sysd.add_particles(x=x, v=v, rho=np.full(n_local, rho0), type=typ)
        sysd.run_program(prog, prog_fields, prog_params, 3)
        sysd.synchronize()
        acc = {k: 0.0 for k in ("move", "cell_list+halo", "balance_of_mass", "find_pressure", "internal_force",
                                "accelerate")}

        def timed(name, fn):
            fn()
            acc[name] += sysd.last_call_ms()

        for _ in range(args.steps):
            timed("move", lambda: sysd.apply(o["move"]))
            timed("cell_list+halo", sysd.create_cell_list)
            timed("balance_of_mass", lambda: sysd.apply(o["bom"]))
            timed("find_pressure", lambda: sysd.apply(o["fp"]))
            timed("internal_force", lambda: sysd.apply(o["force"]))
            timed("accelerate", lambda: sysd.apply(o["acc"]))
            timed("accelerate", lambda: sysd.apply(o["acc"]))
        out["breakdown_ms"] = {k: vv / args.steps for k, vv in acc.items()}
    out["energy"] = sysd.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), (m, c, rho0, *g))[0]
    sysd.close()
    del sysd
    if not e2e:
        return out

    # end to end: upload from pinned host memory, K steps with a per-step all-reduced energy read-back, download
    host_x = torch.from_numpy(x).pin_memory()
    host_v = torch.from_numpy(v).pin_memory()
    host_rho = torch.full((n_local,), rho0, dtype=torch.float64).pin_memory()
    host_typ = torch.from_numpy(np.ascontiguousarray(typ)).pin_memory()
    out_bufs = {}

    def job(s):
        # the NCCL communicator is created once per system outside the timed region (ncclCommInitRank takes ~1 s
        # and a production run pays it once); everything a job moves or computes is inside
        s.resize(n_local)
        for nm, tt in (("x", host_x), ("v", host_v), ("rho", host_rho), ("type", host_typ)):
            s.upload_raw(nm, C.cast(tt.data_ptr(), C.POINTER(C.c_double)), n_local, K["SP_LAYOUT_AOS"])
        e = 0.0
        for _ in range(args.steps):
            slab.wcsph3d_slab_step(s, o)
            e = s.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), (m, c, rho0, *g))[0]
        nl = len(s)
        for nm in ("x", "v", "rho", "P", "_ghost"):
            nc = 3 if nm in ("x", "v") else 1
            key = (nm, nl)
            if key not in out_bufs:
                out_bufs[key] = torch.empty((nl, nc) if nc > 1 else (nl,), dtype=torch.float64).pin_memory()
            s.download_raw(nm, C.cast(out_bufs[key].data_ptr(), C.POINTER(C.c_double)), nl, K["SP_LAYOUT_AOS"])
        s.synchronize()
        return e, nl

    s_warm = make_system()
    s_warm.create_cell_list()
    job(s_warm)
    s_warm.close()
    s_job = make_system()
    # NCCL sets up its peer-to-peer and ring connections lazily at the first send/recv and all-reduce of a new
    # communicator (hundreds of ms): trigger both on the still empty system, outside the timed region
    s_job.create_cell_list()
    s_job.allreduce([0.0])
    s_job.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e_job, nl = job(s_job)
    torch.cuda.synchronize()
    dt_job = time.perf_counter() - t1
    s_job.close()
    tj = torch.tensor([dt_job], dtype=torch.float64)
    dist.all_reduce(tj, op=dist.ReduceOp.MAX)
    h2d = (host_x.numel() + host_v.numel() + host_rho.numel() + host_typ.numel()) * 8
    d2h = nl * 9 * 8
    tot = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    out["e2e"] = {"value": n_tot * args.steps / float(tj[0]), "unit": UNIT, "steps": args.steps,
                  "h2d_bytes_per_step": float(tot[0]) / args.steps,
                  "d2h_bytes_per_step": float(tot[1]) / args.steps + 24 * world, "seconds": float(tj[0]),
                  "what": f"every rank: upload of x,v,rho,type from pinned host memory into a fresh slab system, "
                          f"K = {args.steps} slab steps driven call by call with a per-step all-reduced energy read-back, "
                          f"download of x,v,rho,P,_ghost (NCCL communicator creation excluded); the bulk copies are "
                          f"amortised over K"}
    return out


def box_line(args, UNIT, ClockSampler):
    """BASELINE configs[4] next to the headline workload: the periodic box at per_gpu^3 particles per GPU on the same
    N GPUs (N = 1: one slab exchanging ghosts with itself), so that the per-N lines give its weak-scaling curve."""
    r = run_slab_workload(args, "box", UNIT, ClockSampler, e2e=False, breakdown=False)
    return {k: r[k] for k in ("workload", "value", "ms_per_step", "particles", "particles_per_gpu", "ghosts_per_gpu",
                              "gpu_launches")} | {"unit": UNIT, "steps": args.steps, "warmup": args.warmup}


def run_multi(args, METRIC, UNIT, ClockSampler, peaks, emit):
    dist, rank, world = init_control_plane()
    r = run_slab_workload(args, args.workload, UNIT, ClockSampler)
    other = None
    if args.workload != "box" and not args.no_box:
        other = box_line(args, UNIT, ClockSampler)
    if rank == 0:
        hbm_peak, peak_kind = peaks()
        breakdown = r["breakdown_ms"]
        dominant = max(("internal_force", "balance_of_mass"), key=lambda k: breakdown[k])
        alg = {"internal_force": 120, "balance_of_mass": 72}[dominant]
        kern = {"internal_force": "k_sweep_list<OpInternalForceCached>",
                "balance_of_mass": "k_nbr_build_sweep<OpBalanceOfMassAux> (list build fused with the mass sweep)"}[dominant]
        achieved = alg * r["n_slots"] / (breakdown[dominant] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": r["workload"], "particles": r["particles"],
                       "particles_per_gpu": r["particles_per_gpu"], "particles_per_gpu_all": r["particles_per_gpu_all"],
                       "ghosts_per_gpu": r["ghosts_per_gpu"],
                       "l2": "state >= 1 GB per GPU >> 126 MB L2, no flush needed", "setup_s": r["setup_s"],
                       "parallelism": f"slab{world}", "energy": r["energy"]},
            "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"],
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None, "peak_kind": peak_kind,
                         "alg_bytes_per_particle": alg, "launch_ms": breakdown[dominant],
                         "step_hbm_frac": 736 * r["value"] / world / (hbm_peak * 1e9),
                         "note": "pair sweeps are bound by the L1 data pipe (FP64 gathers), not by HBM"},
            "breakdown_ms": breakdown, "cpu_baseline": None, "box": other,
        }
        emit(line)
    dist.barrier()
    dist.destroy_process_group()
