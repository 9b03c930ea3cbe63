import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices():
    """Number of CUDA devices seen through the C ABI (0 when the library or the driver is missing)"""
    try:
        import ctypes
        from pybinding_b200 import _lib
        count = ctypes.c_int(0)
        if _lib.load().pbk_device_count(ctypes.byref(count)) != 0:
            return 0
        return count.value
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are skipped (not failed) on a box without a CUDA device"""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (run on the B200 box: pytest -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """The reference's own KPM baselines (tests/golden/make_golden.py)"""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_kpm_baselines.npz"))


def max_rel(actual, expected, floor):
    actual, expected = np.asarray(actual), np.asarray(expected)
    return float(np.max(np.abs(actual - expected) / np.maximum(np.abs(expected), floor)))
