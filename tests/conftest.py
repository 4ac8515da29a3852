import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    """gpu-marked tests are skipped only when the product library LOADS and reports no device (or the box has no
    nvidia-smi at all).  A library that is missing or does not load on a GPU box is a failure, not a skip."""
    import shutil
    from defslam_b200 import _capi
    try:
        lib = _capi.load()
    except Exception:
        if shutil.which("nvidia-smi") is None:
            return False       # build container without the library built yet: the CPU tier still runs
        raise
    return lib.defslam_device_count() > 0


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.load()
    return oracle_py


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library.  Missing library on a GPU box is a hard failure, not a skip."""
    from defslam_b200 import _capi
    return _capi.load()
