import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def g_p0():
    return load_golden("p0")


@pytest.fixture(scope="session")
def g_p1():
    return load_golden("p1")


@pytest.fixture(scope="session")
def g_p2():
    return load_golden("p2")


@pytest.fixture(scope="session")
def g_crc():
    return load_golden("crc_ecc")


@pytest.fixture(scope="session")
def g_p4():
    return load_golden("p4")


@pytest.fixture(scope="session")
def g_streams():
    return load_golden("streams")
