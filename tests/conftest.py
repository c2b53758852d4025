import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def _gpu_available():
    try:
        from kover_b200 import _native
        return _native.device_count() > 0
    except Exception:  # noqa: BLE001 (library missing / no driver)
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a GPU skips the `gpu` tests instead of failing in the first one
    (the product has no CPU path to fall back to).  `-m gpu` on a GPU box is unaffected."""
    if not any("gpu" in item.keywords for item in items) or _gpu_available():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (kover_b200 has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
