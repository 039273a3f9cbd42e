import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        from jalla_b200 import _lib
        return _lib.lib.jb_device_count() > 0
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def gpu():
    """Device-resident calls through the C ABI (explicit jb_malloc / jb_memcpy)."""
    from tests.gpu_backend import GpuBackend
    from jalla_b200 import _lib
    rc = _lib.lib.jb_init(-1)
    assert rc == 0, _lib.lib.jb_last_error_string()
    return GpuBackend("device")


@pytest.fixture(scope="session")
def gpu_host():
    """Host pointers handed straight to the C ABI (the library stages them)."""
    from tests.gpu_backend import GpuBackend
    from jalla_b200 import _lib
    assert _lib.lib.jb_init(-1) == 0
    return GpuBackend("host")


@pytest.fixture(scope="session")
def orc():
    from oracle import cpu_ref
    return cpu_ref.oracle()


@pytest.fixture(scope="session")
def blas():
    from oracle import cpu_ref
    return cpu_ref.openblas()


@pytest.fixture
def rng():
    return np.random.default_rng(20261019)
