"""N > 1 host logic on CPU: world_size-2 (and 3) gloo runs of the numpy/oracle emulation of the slab protocol,
plus unit checks of the partition helpers the GPU path shares."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from smoothedparticles_jl_b200 import slab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode,world", [("nonperiodic", 2), ("periodic", 2), ("periodic", 3)])
def test_slab_protocol_emulation_gloo(mode, world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "slab_emulation.py"), mode]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and "EMU-OK" in r.stdout, (r.stdout + r.stderr)[-2000:]


def test_partition_layers_and_owner():
    assert slab.partition_layers(13, 4) == [(0, 4), (4, 7), (7, 10), (10, 13)]
    assert slab.partition_layers(8, 8) == [(i, i + 1) for i in range(8)]
    lay = slab.partition_layers(10, 3)
    assert lay[0][0] == 0 and lay[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(lay, lay[1:]))
    h = 0.01
    z = np.array([-0.0051, 0.0, 0.0099999, 0.01, 0.0949, 0.0999, 0.1])
    own = slab.owner_of(z, h, -1, slab.partition_layers(11, 2))   # cells -1..9 -> layers 0..10
    cells = np.floor(z / h).astype(int) + 1
    assert np.array_equal(own, np.where(cells < 6, 0, np.where(cells < 11, 1, -1)))


def test_balanced_cuts():
    # count-balanced slabs (sp_slab_init_cuts): dense end layers (walls) get thinner slabs, every slab >= 3 layers
    counts = np.array([0, 500, 120, 120] + [120] * 60 + [120, 120, 500, 0])
    for nranks in (2, 3, 4, 8):
        cuts = slab.balanced_cuts(counts, nranks)
        assert cuts[0] == 0 and cuts[-1] == len(counts) and len(cuts) == nranks + 1
        widths = np.diff(cuts)
        assert widths.min() >= 3
        per = [counts[a:b].sum() for a, b in zip(cuts, cuts[1:])]
        assert max(per) - min(per) <= 2 * counts.max()                  # within the granularity of whole layers
        if nranks >= 3:
            assert widths[0] < widths[1] and widths[-1] < widths[-2]    # the wall layers make the end slabs thinner
    assert slab.balanced_cuts(np.ones(24), 8) == list(range(0, 25, 3))
    with pytest.raises(AssertionError):
        slab.balanced_cuts(np.ones(8), 4)


def test_relabelled_axes_give_the_same_dam_break(oracle_lib):
    """configs.collapse3d(slab_axis=0 / 1): the library sees cyclically relabelled coordinates (so a slab decomposition cuts
    along the physical x / y axis); mapped back, the fields after a few steps equal those of the script as it stands (same
    neighbour sets, sums in another visiting order)."""
    import numpy as np
    from oracle.oracle import OracleSystem
    from smoothedparticles_jl_b200 import configs, slab

    base = configs.collapse3d(dr=1.6e-2)
    ref = base.make(OracleSystem)
    for _ in range(4):
        base.step(ref)
    lim_ref = ref.key_lim
    for axis in (0, 1):
        case = configs.collapse3d(dr=1.6e-2, slab_axis=axis)
        perm = case.consts["perm"]
        assert perm[2] == axis and sorted(perm) == [0, 1, 2]
        s = case.make(OracleSystem)
        assert tuple(s.key_lim) == tuple(lim_ref[k] for k in perm)      # the physical axis is now the slowest key axis
        assert slab.slab_axis(s.key_lim) == 2
        for _ in range(4):
            case.step(s)
        assert len(s) == len(ref)
        # particle numbering: both systems keep every particle (nothing leaves in 4 steps), so index i is the same particle
        for name in ("x", "v", "Dv"):
            a = np.empty_like(ref.get(name))
            a[:, perm] = s.get(name)
            b = ref.get(name)
            assert np.max(np.abs(a - b)) <= 1e-11 * max(1.0, np.max(np.abs(b))), name
        for name in ("rho", "P", "type"):
            b = ref.get(name)
            assert np.max(np.abs(s.get(name) - b)) <= 1e-11 * max(1.0, np.max(np.abs(b))), name
