import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "hw_unverified: GPU test of a kernel written after the round's GPU budget ran out: its device code has "
                            "only run under the host emulation (tests/test_cuda_emu.py).  Such tests live in the file that sorts last, "
                            "run as non-strict expected failures (their hardware result is recorded as XPASS / XFAIL without deciding the "
                            "tier's status) and become ordinary tests with B200_RUN_UNVERIFIED=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("B200_RUN_UNVERIFIED") == "1":
        return
    for item in items:
        if "hw_unverified" in item.keywords:
            item.add_marker(pytest.mark.xfail(strict=False, reason="first hardware run of this kernel (host-emulation verified only)"))
            item.add_marker(pytest.mark.timeout(600))


@pytest.fixture(scope="session")
def device():
    """The CUDA device handle every gpu test uses; a missing GPU/library is a hard failure, never a skip."""
    import ffmpeg_b200 as fb
    dev = fb.Device(0)
    dev.set_default()
    yield dev
    dev.close()
