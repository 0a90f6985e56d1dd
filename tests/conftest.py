import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SHIPPED_PKL = os.path.join(GOLDEN, "params_all_split_mutopia_full_aug.pkl")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "reference_numpy_paths.npz"))


@pytest.fixture(scope="session")
def shipped_params():
    from oracle.encoders import load_param_list
    return load_param_list(SHIPPED_PKL)
