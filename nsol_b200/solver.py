"""Abstract solver base of the reference API (``nsol.solver.Solver``).

State and method surface follow nsol/solver.py:21-174: ``x0`` is stored in
scaled units (``x0 / x_scale``, float64), ``get_x()`` / ``get_x0()`` multiply
back, ``run()`` insists on a 1-D ``x0`` (ValueError) and records wall-clock time
as a ``datetime.timedelta``.  Subclasses implement ``_run`` on the GPU.
"""
import datetime
import time
from abc import ABCMeta, abstractmethod

import numpy as np

from nsol_b200 import _lib

_PIN_MIN_ELEMENTS = 1 << 17


def _scaled_host_copy(x0, x_scale):
    """x0 / x_scale as float64 (nsol/solver.py:37).  Large arrays land in page-locked memory when a
    CUDA context already exists, so the later host->device copy runs at full PCIe speed."""
    x0 = np.asarray(x0, dtype=np.float64)
    ctx = _lib.current_context() if x0.size >= _PIN_MIN_ELEMENTS else None
    if ctx is None:
        return np.array(x0, dtype=np.float64) / x_scale
    out = ctx.pinned_empty(x0.shape, np.float64)
    np.divide(x0, x_scale, out=out)
    return out


def _memory_key(a):
    """(address, shape, strides, dtype) of an ndarray's buffer, or None."""
    if not isinstance(a, np.ndarray):
        return None
    return (a.__array_interface__["data"][0], a.shape, a.strides, a.dtype.str)


class Solver(object):
    __metaclass__ = ABCMeta

    def __init__(self, x0, x_scale, verbose):
        self._x_scale = float(x_scale)
        self._x0_src = _memory_key(x0)      # where the caller's x0 lives (lets a solver see that x0 is its observation)
        self._x0 = _scaled_host_copy(x0, self._x_scale)             # nsol/solver.py:37
        self._x = np.array(self._x0)
        self._x_unscaled = None      # get_x() value produced on the device (x * x_scale)
        self._fetch_result = None    # callable downloading the device-resident result
        self._verbose = verbose
        self._computational_time = datetime.timedelta(seconds=0)
        self._observer = None

    def set_x_scale(self, x_scale):
        self._x_scale = x_scale

    def get_x_scale(self):
        return self._x_scale

    def set_verbose(self, verbose):
        self._verbose = verbose

    def get_verbose(self):
        return self._verbose

    def set_x0(self, x0):
        self._x0_src = _memory_key(x0)
        self._x0 = _scaled_host_copy(x0, self._x_scale)
        self._x = np.array(self._x0)
        self._x_unscaled = None
        self._fetch_result = None

    def get_x0(self):
        return np.array(self._x0) * self._x_scale

    def get_x(self):
        """Fresh copy of the solution in original units (nsol/solver.py:117-118).
        After a GPU run the multiplication by x_scale has already been done on the
        device in float64 (bit-identical to the host product)."""
        if self._fetch_result is not None:
            return self._fetch_result()          # fresh array straight from the device
        if self._x_unscaled is not None:
            return np.array(self._x_unscaled)
        return np.array(self._x) * self._x_scale

    def get_computational_time(self):
        return self._computational_time

    def set_observer(self, observer):
        self._observer = observer

    def run(self):
        if self._x0.ndim != 1:
            raise ValueError("Initial value x0 must be a 1D array")   # nsol/solver.py:149-150
        t0 = time.time()
        self._run()
        self._computational_time = datetime.timedelta(seconds=time.time() - t0)
        if self._verbose:
            print("Required computational time: %s" % (self.get_computational_time()))
        if self._observer is not None:
            self._observer.set_computational_time(self.get_computational_time())

    # results coming back from the device
    def _set_result(self, x_unscaled):
        self._fetch_result = None
        self._x_unscaled = x_unscaled
        self._x = _LazyScaled(lambda: x_unscaled, self._x_scale)

    def _set_device_result(self, fetch):
        """The solution stays on the device; every get_x() downloads a fresh float64 array
        (x * x_scale evaluated on the device), like the reference returns a fresh array."""
        self._x_unscaled = None
        self._fetch_result = fetch
        self._x = _LazyScaled(fetch, self._x_scale)

    @abstractmethod
    def _run(self):
        pass

    @abstractmethod
    def print_statistics(self):
        pass


class _LazyScaled(object):
    """``solver._x`` after a device run: materialises x_unscaled / x_scale on demand
    (only the diagnostic cost getters of LinearSolver read it)."""

    def __init__(self, fetch, x_scale):
        self._fetch, self._s = fetch, x_scale

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._fetch() / self._s, dtype=dtype)
