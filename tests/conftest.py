"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"` runs the oracle-vs-golden checks, host logic, C-ABI symbol
checks and gloo multi-process tests on CPU; `-m gpu` runs the CUDA parity
tests through the C-ABI on a B200.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


class Golden(object):
    """Lazy access to tests/golden/*.npz (outputs of the unmodified reference)."""

    def __init__(self):
        self._files = {}
        with open(os.path.join(GOLDEN, "manifest.json")) as fh:
            self.manifest = json.load(fh)

    def __call__(self, group, key):
        if group not in self._files:
            self._files[group] = np.load(os.path.join(GOLDEN, group + ".npz"))
        return self._files[group][key]


@pytest.fixture(scope="session")
def golden():
    return Golden()


def rel_max(a, b):
    """relative max-abs error  max|a-b| / max|b|  (north_star's parity measure)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
