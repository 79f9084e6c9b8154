import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRID3D = os.path.join(ROOT, "grids", "box_3D_elongated.npz")
GRID2D = os.path.join(ROOT, "grids", "refined.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gpu_backend():
    from admm_optim_b200 import ug4
    return ug4.Backend(device=0)
