import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """The reference's own KPM baselines (tests/golden/make_golden.py)"""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_kpm_baselines.npz"))


def max_rel(actual, expected, floor):
    actual, expected = np.asarray(actual), np.asarray(expected)
    return float(np.max(np.abs(actual - expected) / np.maximum(np.abs(expected), floor)))
