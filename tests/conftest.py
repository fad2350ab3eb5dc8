import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the CPU suite checks that the C-ABI library loads and exports every declared symbol:
    # build it (nvcc cross-compiles without a GPU) if this is a fresh checkout
    lib = os.path.join(ROOT, "fastbox_b200", "libfastbox_b200.so")
    if not os.path.isfile(lib):
        try:
            import __graft_entry__
            __graft_entry__.build()
        except Exception as exc:      # pragma: no cover
            print("WARNING: could not build libfastbox_b200.so: %s" % exc)


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
