import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    """Asks the library itself (cudaGetDeviceCount behind porla_device_count): the product needs only the CUDA runtime,
    so a missing or CPU-only torch must not turn the parity tests into skips."""
    try:
        from porla_b200 import lib
        return int(lib.load().porla_device_count()) > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    if os.environ.get("PORLA_REQUIRE_GPU") == "1" and any("gpu" in item.keywords for item in items):
        raise pytest.UsageError("PORLA_REQUIRE_GPU=1 but libmultiexp.so sees no CUDA device (or is not built)")
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
