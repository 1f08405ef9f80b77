"""CPU emulation of the slab protocol of csrc/sp_slab.cu, one process per rank over gloo: the same layer partition, the
same ONE-ROUND exchange rule and periodic shifts, with numpy for the packing and the CPU oracle as each rank's local
engine, over several steps with particles migrating between ranks.

The rule being emulated (sp_slab.cu, header comment):
  * local window = owned layers [c0, c1) + W = 2 ghost layers per side;
  * every rebuild: old ghosts are dropped; every OWNED particle whose current layer is < c0 + 2 goes into the message to
    the lower neighbour, every one with layer >= c1 - 2 into the message to the upper neighbour (ghost copies and
    migrants alike); the sender keeps its copies; ownership afterwards is a function of the position alone;
  * consequence checked here: a rank's two boundary layers and its neighbour's two ghost layers hold the same particles.
Rank 0 checks the merged densities of the owned particles against a single-domain oracle run at every step."""
import os
import sys

import numpy as np
import torch  # noqa: F401  (torch.distributed needs it initialised)
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smoothedparticles_jl_b200 as sp  # noqa: E402,F401
from smoothedparticles_jl_b200 import geometry as geo, operators as ops, slab  # noqa: E402
from oracle.oracle import OracleSystem  # noqa: E402

W = 2


def exchange(send_dn, send_up, below, above):
    """What I send down arrives at my lower neighbour from above (same call order as slab_exchange_payload)."""
    width = send_dn.shape[1]
    out = {"lo": np.zeros((0, width)), "hi": np.zeros((0, width))}
    table = [None] * dist.get_world_size()
    dist.all_gather_object(table, {"dn": (below, send_dn), "up": (above, send_up)})
    me = dist.get_rank()
    for t in table:
        if t["dn"][0] == me:   # sent DOWN to me: arrives from above
            out["hi"] = t["dn"][1]
        if t["up"][0] == me:   # sent UP to me: arrives from below
            out["lo"] = t["up"][1]
    return out["lo"], out["hi"]


def main():
    periodic = sys.argv[1] == "periodic"
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, nz = 8, 7, 24
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
    assert min(b - a for a, b in layers) >= 3, "three owned layers per rank (sp_slab_init)"
    assert slab.slab_axis((5, 5, glim)) == 2 and slab.slab_axis((5, 5, 1)) == 1
    owner = slab.owner_of(x[:, 2], h, gphase, layers)
    assert owner.min() >= 0
    c0, c1 = layers[rank]
    below, above = rank - 1, rank + 1
    if periodic:
        below, above = (rank - 1) % world, (rank + 1) % world
    elif above >= world:
        above = -1
    # per-particle drift along z, up to 0.45 cells per step in either direction (a function of the global id)
    gid_all = np.arange(n)
    vz_all = 0.45 * h * np.sin(0.37 * gid_all + 0.1)
    m = 1000.0 * dr ** 3
    mine = owner == rank
    # local state: rows [x, y, z, gid], ghost flag
    loc = np.column_stack([x[mine], gid_all[mine].astype(float)])
    ghost = np.zeros(len(loc), dtype=bool)
    xg = x.copy()          # the single-domain reference (rank 0 uses it)
    alive_g = np.ones(n, dtype=bool)
    ok = True
    worst = 0.0
    for step in range(7):
        # ---- move (ghosts too: they are dropped anyway)
        if step > 0:
            loc[:, 2] += vz_all[loc[:, 3].astype(int)]
            xg[:, 2] += vz_all
            if periodic:
                xg[:, 2] = np.where(xg[:, 2] >= Lz, xg[:, 2] - Lz, np.where(xg[:, 2] < 0.0, xg[:, 2] + Lz, xg[:, 2]))
            else:
                alive_g &= (xg[:, 2] >= dom.lo[2]) & (xg[:, 2] <= dom.hi[2])
        # ---- the one-round exchange
        loc, ghost = loc[~ghost], ghost[~ghost]                       # old ghosts die
        lay = np.floor(loc[:, 2] / h).astype(np.int64) - gphase
        send_dn = loc[lay < c0 + W] if below >= 0 else loc[:0]
        send_up = loc[lay >= c1 - W] if above >= 0 else loc[:0]
        g_lo, g_hi = exchange(send_dn, send_up, below, above)
        if periodic and rank == 0:
            g_lo = g_lo - np.array([0, 0, Lz, 0])
        if periodic and rank == world - 1:
            g_hi = g_hi + np.array([0, 0, Lz, 0])
        loc = np.concatenate([loc, g_lo, g_hi])
        lay = np.floor(loc[:, 2] / h).astype(np.int64) - gphase
        keep = (lay >= c0 - W) & (lay < c1 + W)
        if not periodic:
            keep &= (loc[:, 2] >= dom.lo[2]) & (loc[:, 2] <= dom.hi[2])
        loc, lay = loc[keep], lay[keep]
        ghost = (lay < c0) | (lay >= c1)
        ids = loc[:, 3].astype(int)
        assert len(np.unique(ids)) == len(ids), "a particle is held twice by one rank"
        # ---- density of the owned particles on the local window
        zlo = (gphase + c0 - W) * h + 1e-9 * h
        zhi = (gphase + c1 + W) * h - 1e-9 * h
        ldom = geo.Box(dom.lo[0], dom.lo[1], zlo, dom.hi[0], dom.hi[1], zhi)
        ora = OracleSystem({"rho": 1}, ldom, h)
        ora.add_particles(x=np.ascontiguousarray(loc[:, :3]))
        ora.create_cell_list()
        assert len(ora) == len(loc), "a particle fell outside the local window"
        ora.apply(ops.density_sum("wendland3", m, h), self_=True)
        rho = ora.get("rho")
        # ---- gather: owned (gid, rho), and the boundary / ghost id sets for the invariant
        own = ~ghost
        zone = {"bnd_dn": set(ids[own & (lay < c0 + W)]), "bnd_up": set(ids[own & (lay >= c1 - W)]),
                "gh_lo": set(ids[lay < c0]), "gh_hi": set(ids[lay >= c1]), "below": below, "above": above}
        out = [None] * world
        dist.all_gather_object(out, (ids[own], rho[own], zone))
        if rank == 0:
            allg = np.concatenate([o[0] for o in out])
            allr = np.concatenate([o[1] for o in out])
            order = np.argsort(allg)
            allg, allr = allg[order], allr[order]
            xa = xg[alive_g]
            if periodic:
                lo_img = xa[xa[:, 2] >= Lz - W * h] - np.array([0, 0, Lz])
                hi_img = xa[xa[:, 2] < W * h] + np.array([0, 0, Lz])
                xall = np.concatenate([xa, lo_img, hi_img])
                refdom = geo.Box(dom.lo[0], dom.lo[1], -W * h, dom.hi[0], dom.hi[1], Lz + W * h)
            else:
                xall, refdom = xa, dom
            ref = OracleSystem({"rho": 1}, refdom, h)
            ref.add_particles(x=xall)
            ref.create_cell_list()
            ref.apply(ops.density_sum("wendland3", m, h), self_=True)
            rr = ref.get("rho")[: len(xa)]
            same_set = np.array_equal(allg, gid_all[alive_g])
            err = np.max(np.abs(allr - rr)) / np.max(rr) if same_set else np.inf
            worst = max(worst, err)
            ok = ok and same_set and err <= 1e-12
            # invariant behind the zero-copy halo refresh: boundary layers of the owner == ghost layers of the holder
            for r, (_, _, z) in enumerate(out):
                if z["below"] >= 0:
                    ok = ok and z["bnd_dn"] == out[z["below"]][2]["gh_hi"]
                if z["above"] >= 0:
                    ok = ok and z["bnd_up"] == out[z["above"]][2]["gh_lo"]
    if rank == 0:
        print(("EMU-OK" if ok else "EMU-FAIL"), sys.argv[1], world, worst, flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
