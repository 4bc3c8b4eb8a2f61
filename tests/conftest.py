import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """A persistent kernel that waits for a box nobody issues would hang the whole run: every GPU test gets a
    time limit (pytest-timeout, thread method: the process is ended, the run reports the test)."""
    try:
        import pytest_timeout  # noqa: F401
    except ImportError:
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(180, method="thread"))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("test marked gpu but no CUDA device is visible")
    from multicamera_stitching_b200 import _cabi
    _cabi.load()  # fail loudly if the extension is missing
    return torch.device("cuda", 0)
