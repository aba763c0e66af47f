import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import fclibs
    return fclibs.oracle()


@pytest.fixture(scope="session")
def ref():
    import fclibs
    lib = fclibs.reference()
    if lib is None:
        pytest.skip("oracle/_ref/libfcref.so not built (needs /root/reference)")
    return lib


@pytest.fixture(scope="session")
def gpu():
    """The product library on a real device.  Fails (not skips) if the native library is missing."""
    import torch
    import fclibs
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    api = fclibs.product()
    assert api.device_count() >= 1, api.last_error()
    return api
