"""Adaptive CFL time step (BASELINE north_star; an extension, the reference's examples use fixed steps): the device
max-speed reduction against the oracle and against numpy."""
import numpy as np
import pytest

import smoothedparticles_jl_b200 as sp
from smoothedparticles_jl_b200 import ParticleSystem, configs
from oracle.oracle import OracleSystem

pytestmark = pytest.mark.gpu
K = sp.K


def test_max_speed_reduction_and_cfl_step():
    case = configs.lattice_box(40, jitter=0.1, dr=5e-3)   # 64 000 particles, shuffled order
    c = case.consts
    dev, ora = case.make(ParticleSystem), case.make(OracleSystem)
    v = case.init["v"]
    want = np.max(np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]))
    for s in (dev, ora):
        assert s.reduce(K["SP_RED_MAX_SPEED"], ("v",), (), nout=1)[0] == want      # a maximum: exact
    for _ in range(5):
        case.step(dev)
        case.step(ora)
    dt_d, dt_o = sp.cfl_time_step(dev, 0.1, case.h, c["c"]), sp.cfl_time_step(ora, 0.1, case.h, c["c"])
    assert abs(dt_d - dt_o) <= 1e-10 * dt_o
    assert 0.0 < dt_d < 0.1 * case.h / c["c"]
    empty = ParticleSystem({"v": 3}, case.domain, case.h)
    assert empty.reduce(K["SP_RED_MAX_SPEED"], ("v",), (), nout=1)[0] == 0.0
