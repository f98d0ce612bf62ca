import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "hw_unverified: GPU test of a kernel written after the round's GPU budget ran out; it has "
                            "never run on hardware, so it only runs with B200_RUN_UNVERIFIED=1 (first thing next round)")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("B200_RUN_UNVERIFIED") == "1":
        return
    skip = pytest.mark.skip(reason="kernel not yet run on hardware (written without GPU access); set B200_RUN_UNVERIFIED=1")
    for item in items:
        if "hw_unverified" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def device():
    """The CUDA device handle every gpu test uses; a missing GPU/library is a hard failure, never a skip."""
    import ffmpeg_b200 as fb
    dev = fb.Device(0)
    dev.set_default()
    yield dev
    dev.close()
