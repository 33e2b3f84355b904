import os
import sys

import pytest

os.environ.setdefault("OPENCV_LOG_LEVEL", "SILENT")       # libtiff's "unknown GeoTIFF tag" chatter through OpenCV
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle / synthetic helpers are compiled (gcc only; the CUDA library is prebuilt)."""
    import oracle
    import synthetic
    oracle.build()
    synthetic.build()
