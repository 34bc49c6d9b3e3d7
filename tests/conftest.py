import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_c1():
    return np.load(os.path.join(ROOT, "tests", "golden", "c1_reference_scene.npz"))


@pytest.fixture(scope="session")
def golden_wind():
    return np.load(os.path.join(ROOT, "tests", "golden", "wind_n32.npz"))


def full_state(posvel, corr=None):
    """[S,2,N,3] pos+vel xyz (+ optional corr [S,N,3]) -> Strand[S] AoS [S,3,N,4]."""
    S, _, N, _ = posvel.shape
    st = np.zeros((S, 3, N, 4), np.float32)
    st[:, 0:2, :, :3] = posvel
    st[:, 0, :, 3] = 1.0
    if corr is not None:
        st[:, 2, :, :3] = corr
    return st
