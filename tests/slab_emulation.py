"""CPU emulation of the slab protocol of csrc/sp_slab.cu, one process per rank over gloo: the same layer
partition, ownership rule, boundary-layer selection and periodic shifts, with numpy for the packing and the CPU
oracle as each rank's local engine.  Rank 0 checks the merged result against a single-domain oracle run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smoothedparticles_jl_b200 as sp  # noqa: E402
from smoothedparticles_jl_b200 import geometry as geo, operators as ops, slab  # noqa: E402
from oracle.oracle import OracleSystem  # noqa: E402


def exchange(send_dn, send_up, below, above):
    """What I send down arrives at my lower neighbour from above (same call order as slab_exchange_payload)."""
    out = {"lo": np.zeros((0, send_dn.shape[1])), "hi": np.zeros((0, send_dn.shape[1]))}
    reqs = []
    if below >= 0:
        reqs.append(("s", below, send_dn))
    if above >= 0:
        reqs.append(("s", above, send_up))
    objs_from = {}
    # gloo object exchange keeps the emulation short: every rank publishes what it sends to whom
    table = [None] * dist.get_world_size()
    dist.all_gather_object(table, {"dn": (below, send_dn), "up": (above, send_up)})
    me = dist.get_rank()
    for r, t in enumerate(table):
        if t["dn"][0] == me:   # r sent DOWN to me: arrives from above
            out["hi"] = t["dn"][1]
        if t["up"][0] == me:   # r sent UP to me: arrives from below
            out["lo"] = t["up"][1]
    return out["lo"], out["hi"]


def main():
    periodic = sys.argv[1] == "periodic"
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, nz = 10, 9, 16
    dr = 5e-3
    h = 2 * dr
    rng = np.random.default_rng(5)
    I, J, Kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    x = np.stack([I.ravel(), J.ravel(), Kk.ravel()], 1) * dr + rng.uniform(0.05 * dr, 0.95 * dr, size=(nx * ny * nz, 3))
    n = len(x)
    Lz = nz * dr
    dom = geo.Box(-h, -h, 0.0, nx * dr + h, ny * dr + h, Lz * (1 - 1e-12))
    gphase = int(np.floor(dom.lo[2] / h))
    glim = int(np.floor(dom.hi[2] / h)) - gphase + 1
    layers = slab.partition_layers(glim, world)
    assert slab.slab_axis((5, 5, glim)) == 2 and slab.slab_axis((5, 5, 1)) == 1
    owner = slab.owner_of(x[:, 2], h, gphase, layers)
    assert owner.min() >= 0 and np.array_equal(np.bincount(owner, minlength=world) > 0, np.ones(world, bool))
    c0, c1 = layers[rank]
    mine = owner == rank
    gid = np.arange(n)[mine]
    xo = x[mine]
    cell = np.floor(xo[:, 2] / h).astype(np.int64) - gphase
    below, above = rank - 1, rank + 1
    if periodic:
        below, above = (rank - 1) % world, (rank + 1) % world
    elif above >= world:
        above = -1
    send_dn = xo[cell == c0] if below >= 0 else xo[:0]
    send_up = xo[cell == c1 - 1] if above >= 0 else xo[:0]
    g_lo, g_hi = exchange(send_dn, send_up, below, above)
    if periodic and rank == 0:
        g_lo = g_lo - np.array([0, 0, Lz])
    if periodic and rank == world - 1:
        g_hi = g_hi + np.array([0, 0, Lz])
    local = np.concatenate([xo, g_lo, g_hi])
    # local window: owned layers + one ghost layer per side (floats chosen strictly inside the end cells)
    zlo = (gphase + c0 - 1) * h + 1e-9 * h
    zhi = (gphase + c1 + 1) * h - 1e-9 * h
    ldom = geo.Box(dom.lo[0], dom.lo[1], zlo, dom.hi[0], dom.hi[1], zhi)
    m = 1000.0 * dr ** 3
    ora = OracleSystem({"rho": 1}, ldom, h)
    ora.add_particles(x=local)
    ora.create_cell_list()
    assert len(ora) == len(local), "a ghost fell outside the local window"
    ora.apply(ops.density_sum("wendland3", m, h), self_=True)
    rho = ora.get("rho")[: len(xo)]
    out = [None] * world
    dist.all_gather_object(out, (gid, rho))
    ok = True
    if rank == 0:
        allg = np.concatenate([o[0] for o in out])
        allr = np.concatenate([o[1] for o in out])[np.argsort(allg)]
        if periodic:
            lo_img = x[x[:, 2] >= Lz - h] - np.array([0, 0, Lz])
            hi_img = x[x[:, 2] < h] + np.array([0, 0, Lz])
            xa = np.concatenate([x, lo_img, hi_img])
            refdom = geo.Box(dom.lo[0], dom.lo[1], -h, dom.hi[0], dom.hi[1], Lz + h)
        else:
            xa, refdom = x, dom
        ref = OracleSystem({"rho": 1}, refdom, h)
        ref.add_particles(x=xa)
        ref.create_cell_list()
        ref.apply(ops.density_sum("wendland3", m, h), self_=True)
        err = np.max(np.abs(allr - ref.get("rho")[:n])) / np.max(ref.get("rho"))
        ok = len(allg) == n and err <= 1e-12
        print(("EMU-OK" if ok else "EMU-FAIL"), sys.argv[1], world, err, flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
