import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running exhaustive check")


@pytest.fixture(scope="session")
def goldens():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "cv2_orb_goldens.npz"))
