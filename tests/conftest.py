import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        from fastbox_b200 import _lib
        _lib.load()
        _lib.device_info(0)
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """GPU tests fail loudly (not skip) when the library or device is missing."""
    from fastbox_b200 import _lib
    _lib.load()
    name, sms, mem = _lib.device_info(0)
    return dict(name=name, sms=sms, mem=mem)


def pytest_collection_modifyitems(config, items):
    # when no -m expression is given on a machine without a GPU, skip gpu tests
    if config.getoption("-m"):
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
