"""Worker of the multi-process slab tests: launched once per rank (torch.distributed.run or the test itself),
gloo for the control plane, NCCL inside the library for the data plane.  Rank 0 checks the gathered result
against the single-domain CPU oracle and exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import smoothedparticles_jl_b200 as sp  # noqa: E402
from smoothedparticles_jl_b200 import configs, geometry as geo, operators as ops, slab  # noqa: E402


def gather_by_gid(sysd, names, rank, world):
    """Owned particles of every rank, merged and ordered by the global id field."""
    mask = sysd.owned_mask()
    payload = {nm: sysd.get(nm)[mask] for nm in names + ["gid"]}
    out = [None] * world
    dist.all_gather_object(out, payload)
    merged = {nm: np.concatenate([o[nm] for o in out]) for nm in names + ["gid"]}
    order = np.argsort(merged["gid"], kind="stable")
    return {nm: merged[nm][order] for nm in merged}


def main():
    case_name = sys.argv[1]
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = [slab.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ndev = torch.cuda.device_count()
    device = local % max(ndev, 1)
    K = sp.K
    ok = True
    msg = ""
    if case_name in ("box_nonperiodic", "box_steps", "box_program"):
        case = configs.lattice_box((20, 18, 26), jitter=0.15, dr=5e-3, seed=7)
        c = case.consts
        n = case.n
        fields = dict(case.fields)
        fields["gid"] = 1
        sysd = slab.SlabSystem(fields, case.domain, case.h, rank, world, ids[0], periodic=False, device=device)
        mine = sysd.owns(case.init["x"])
        init = {k: v[mine] for k, v in case.init.items()}
        init["gid"] = np.arange(n, dtype=float)[mine]
        sysd.add_particles(**init)
        o = case.ops
        names = ["x", "v", "rho", "P", "Dv"]
        nsteps = 1 if case_name == "box_nonperiodic" else 12
        if case_name == "box_program":
            # the same 12 steps issued from inside the library (sp_run_program on a slab system)
            sysd.run_program(case.program, case.program_fields, case.program_params, nsteps)
        else:
            for _ in range(nsteps):
                slab.wcsph3d_slab_step(sysd, o)
        tot = sysd.allreduce([sysd.n_owned])[0]
        res = gather_by_gid(sysd, names, rank, world)
        E = sysd.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), (c["m"], c["c"], c["rho0"], *c["g"]))[0]
        # adaptive CFL step: the max-speed reduction is all-reduced (NCCL max), every rank gets the same dt
        dt_cfl = sp.cfl_time_step(sysd, 0.1, case.h, c["c"])
        dts = [None] * world
        dist.all_gather_object(dts, dt_cfl)
        ok = ok and all(d == dts[0] for d in dts)
        if rank == 0:
            from oracle.oracle import OracleSystem
            ora = case.make(OracleSystem)
            for _ in range(nsteps):
                case.step(ora)
            Eo = ora.reduce(K["SP_RED_ENERGY_WCSPH"], ("x", "v", "rho"), (c["m"], c["c"], c["rho0"], *c["g"]))[0]
            ok = ok and int(tot) == len(ora) == len(res["gid"])
            msg += f"count {int(tot)} vs {len(ora)}; "
            tol = 1e-10 if nsteps == 1 else 1e-9
            for nm in names:
                a, b = res[nm], ora.get(nm)
                err = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
                msg += f"{nm}:{err:.2e} "
                ok = ok and err <= tol
            ok = ok and abs(E - Eo) <= 1e-9 * abs(Eo)
            msg += f"E {E:.12e} vs {Eo:.12e}"
            dt_o = sp.cfl_time_step(ora, 0.1, case.h, c["c"])
            ok = ok and abs(dt_cfl - dt_o) <= 1e-9 * dt_o and dt_o < 0.1 * case.h / c["c"]
            msg += f" dt_cfl {dt_cfl:.12e} vs {dt_o:.12e}"
    elif case_name == "box_lists":
        # SURVEY 8(e) "parity under sharding": neighbour sets per GLOBAL id, bit-exact, after 12 steps with migration.
        # The oracle gets the device positions (bit-exactness is a statement about identical inputs).
        case = configs.lattice_box((20, 18, 26), jitter=0.15, dr=5e-3, seed=7)
        n = case.n
        fields = dict(case.fields)
        fields["gid"] = 1
        sysd = slab.SlabSystem(fields, case.domain, case.h, rank, world, ids[0], periodic=False, device=device)
        mine = sysd.owns(case.init["x"])
        init = {k: v[mine] for k, v in case.init.items()}
        init["gid"] = np.arange(n, dtype=float)[mine]
        sysd.add_particles(**init)
        for _ in range(12):
            slab.wcsph3d_slab_step(sysd, case.ops)
        off, idx = sysd.neighbour_lists()                 # local 1-based indices, device order
        gid_loc = sysd.get("gid").astype(np.int64)
        own = sysd.owned_mask()
        nb = {}
        for i in np.nonzero(own)[0]:
            nb[int(gid_loc[i])] = np.sort(gid_loc[idx[off[i]:off[i + 1]] - 1])
        res = gather_by_gid(sysd, ["x"], rank, world)
        allnb = [None] * world
        dist.all_gather_object(allnb, nb)
        if rank == 0:
            from oracle.oracle import OracleSystem
            ora = OracleSystem({"gid": 1}, case.domain, case.h)
            ora.add_particles(x=res["x"], gid=res["gid"])
            ora.create_cell_list()
            oo, oi = ora.neighbour_lists()
            merged = {}
            for d in allnb:
                merged.update(d)
            ok = len(merged) == n == len(ora)
            bad = 0
            for g in range(n):
                ref = np.sort(oi[oo[g]:oo[g + 1]] - 1)
                if not np.array_equal(ref, merged.get(g, np.array([-1]))):
                    bad += 1
            ok = ok and bad == 0
            msg = f"neighbour sets by global id: {n} particles, {int(oo[-1])} pairs, {bad} mismatching lists"
    elif case_name == "isph_cg":
        # ISPH on slabs (2-D: the slab axis is y): pre-solve operators, matrix-free CG with NCCL all-reduced dot products
        # and a halo refresh of the search vector per iteration, against the oracle's assembled matrix + CG
        # (collapse_dry_implicit.jl:218-233)
        case = configs.collapse_dry_implicit(dr=2.0e-2)
        rng = np.random.default_rng(2)
        case.init["v"] = rng.uniform(-1, 1, size=(case.n, 3)) * np.array([1.0, 1.0, 0.0])
        n = case.n
        o = case.ops
        fields = dict(case.fields)
        fields["gid"] = 1
        sysd = slab.SlabSystem(fields, case.domain, case.h, rank, world, ids[0], periodic=False, device=device)
        mine = sysd.owns(case.init["x"])
        init = {k: v[mine] for k, v in case.init.items()}
        init["gid"] = np.arange(n, dtype=float)[mine]
        sysd.add_particles(**init)
        case.prologue(sysd)
        sysd.apply(o["init"])
        sysd.create_cell_list()
        sysd.apply(o["visc"])
        sysd.apply(o["dll"])
        sysd.apply(o["b"])
        it_d, res_d = sysd.poisson_cg(o["A"], "b", "P")
        sysd.apply(o["force"])
        sysd.apply(o["acc"])
        res = gather_by_gid(sysd, ["b", "P", "v", "L", "lambda"], rank, world)
        if rank == 0:
            from oracle.oracle import OracleSystem
            ora = case.make(OracleSystem)
            case.prologue(ora)
            ora.apply(o["init"])
            ora.create_cell_list()
            ora.apply(o["visc"])
            ora.apply(o["dll"])
            ora.apply(o["b"])
            I, J, V = ora.assemble_matrix(o["A"])
            x_ref, it_o, res_o = ora.cg(I, J, V, ora.get("b"))
            ora.set("P", x_ref)
            ora.apply(o["force"])
            ora.apply(o["acc"])
            ok = len(res["gid"]) == n == len(ora)
            errs = {}
            for nm, tol in (("L", 1e-10), ("lambda", 1e-10), ("b", 1e-10), ("P", 1e-5), ("v", 1e-6)):
                a, b_ = res[nm], ora.get(nm)
                errs[nm] = np.max(np.abs(a - b_)) / max(np.max(np.abs(b_)), 1e-300)
                ok = ok and errs[nm] <= tol
            ok = ok and abs(it_d - it_o) <= max(3, it_o // 20)
            ok = ok and res_d <= np.sqrt(np.finfo(float).eps) * np.linalg.norm(ora.get("b"))
            msg = f"iters {it_d} vs {it_o}; " + " ".join(f"{k}:{v:.2e}" for k, v in errs.items())
    elif case_name == "box_periodic":
        # periodic along z: compare with an oracle run on the same particles plus explicit periodic images
        nx, ny, nz = 14, 12, 24
        dr = 5e-3
        h = 2 * dr
        rng = np.random.default_rng(5)
        I, J, Kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        x = np.stack([I.ravel(), J.ravel(), Kk.ravel()], 1) * dr + rng.uniform(0.05 * dr, 0.95 * dr, size=(nx * ny * nz, 3))
        n = len(x)
        Lz = nz * dr                                   # = 12 cell layers of h
        dom = geo.Box(-h, -h, 0.0, nx * dr + h, ny * dr + h, Lz * (1 - 1e-12))
        sysd = slab.SlabSystem({"rho": 1, "gid": 1}, dom, h, rank, world, ids[0], periodic=True, device=device)
        mine = sysd.owns(x)
        sysd.add_particles(x=x[mine], gid=np.arange(n, dtype=float)[mine])
        sysd.create_cell_list()
        m = 1000.0 * dr ** 3
        sysd.apply(ops.density_sum("wendland3", m, h), self_=True)
        res = gather_by_gid(sysd, ["rho", "x"], rank, world)
        if rank == 0:
            from oracle.oracle import OracleSystem
            lo_img = x[x[:, 2] >= Lz - h] - np.array([0, 0, Lz])
            hi_img = x[x[:, 2] < h] + np.array([0, 0, Lz])
            xa = np.concatenate([x, lo_img, hi_img])
            domo = geo.Box(-h, -h, -h, nx * dr + h, ny * dr + h, Lz + h)
            ora = OracleSystem({"rho": 1}, domo, h)
            ora.add_particles(x=xa)
            ora.create_cell_list()
            ora.apply(ops.density_sum("wendland3", m, h), self_=True)
            b = ora.get("rho")[:n]
            err = np.max(np.abs(res["rho"] - b)) / np.max(np.abs(b))
            ok = err <= 1e-10 and np.array_equal(res["x"], x)
            msg = f"periodic rho err {err:.2e}"
    else:
        ok, msg = False, "unknown case"
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    if rank == 0:
        print(("SLAB-OK " if ok else "SLAB-FAIL ") + case_name + " world=%d " % world + msg, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(3)
