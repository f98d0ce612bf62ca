import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def device():
    """The CUDA device handle every gpu test uses; a missing GPU/library is a hard failure, never a skip."""
    import ffmpeg_b200 as fb
    dev = fb.Device(0)
    dev.set_default()
    yield dev
    dev.close()
