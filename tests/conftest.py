import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def example_prefix(tmp_path_factory):
    import orclib
    d = tmp_path_factory.mktemp("example_index")
    return orclib.materialise_example_index(str(d))


@pytest.fixture(scope="session")
def golden_read():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "example_read.npz"))
