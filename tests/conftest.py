import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def psf_npz_path():
    """The reference's only real numeric fixture (sample_data/psf.npz), committed as a
    copy under tests/golden/ because /root/reference does not exist on the GPU box."""
    return os.path.join(ROOT, "tests", "golden", "psf.npz")
