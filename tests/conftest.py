import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

    class G(object):
        def __getitem__(self, name):
            return np.load(os.path.join(gdir, name + ".npz"))

        def has(self, name):
            return os.path.isfile(os.path.join(gdir, name + ".npz"))
    return G()
