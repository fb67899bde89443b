import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.load_library()
    return oracle


@pytest.fixture()
def cpu(oracle_mod):
    o = oracle_mod.Oracle()
    yield o
    o.close()


@pytest.fixture()
def gpu():
    from vofod_b200 import capi
    g = capi.Vofod(0)  # raises loudly when libvofod_cuda.so or the device is missing: there is no CPU fallback
    yield g
    g.close()
