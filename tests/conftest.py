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


# ---- bounds checking of the CUDA kernels (stands in for compute-sanitizer memcheck, which is closed on the GPU pool) ----------
# NSOL_DEBUG_GUARD=1 python -m pytest tests -m gpu : every plan array of every GPU test sits between two NaN-filled 64 KiB guard
# bands (nsol_set_tuning "debug_guard").  An out-of-bounds READ drags NaN into a result the parity assertions look at; an
# out-of-bounds WRITE is counted by nsol_debug_guard_check after each test.
@pytest.fixture(autouse=True)
def _guard_bands(request):
    if os.environ.get("NSOL_DEBUG_GUARD", "0") != "1" or request.node.get_closest_marker("gpu") is None:
        yield
        return
    from nsol_b200 import _lib
    ctx = _lib.context()
    ctx.set_tuning("debug_guard", 1)
    yield
    bad, _ = ctx.guard_check()
    assert bad == 0, "%s: %d guard bytes overwritten (out-of-bounds write)" % (request.node.name, bad)
