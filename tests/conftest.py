import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def _gpu_available():
    try:
        import ctypes as C

        import smoothedparticles_jl_b200 as sp
        lib = sp.abi.load()
        n = C.c_int32()
        return lib.sp_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu tests must FAIL (not skip) on a box whose GPU/extension is unusable only when explicitly selected;
    # when not selected they are deselected by the marker expression anyway.
    pass


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def sp():
    import smoothedparticles_jl_b200 as sp
    return sp
