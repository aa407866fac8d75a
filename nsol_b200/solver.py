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


class Solver(object):
    __metaclass__ = ABCMeta

    def __init__(self, x0, x_scale, verbose):
        self._x_scale = float(x_scale)
        self._x0 = np.array(x0, dtype=np.float64) / self._x_scale   # nsol/solver.py:37
        self._x = np.array(self._x0)
        self._x_unscaled = None      # get_x() value produced on the device (x * x_scale)
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
        self._x0 = np.array(x0, dtype=np.float64) / self._x_scale
        self._x = np.array(self._x0)
        self._x_unscaled = None

    def get_x0(self):
        return np.array(self._x0) * self._x_scale

    def get_x(self):
        """Fresh copy of the solution in original units (nsol/solver.py:117-118).
        After a GPU run the multiplication by x_scale has already been done on the
        device in float64 (bit-identical to the host product)."""
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
        self._x_unscaled = x_unscaled
        self._x = _LazyScaled(x_unscaled, self._x_scale)

    @abstractmethod
    def _run(self):
        pass

    @abstractmethod
    def print_statistics(self):
        pass


class _LazyScaled(object):
    """``solver._x`` after a device run: materialises x_unscaled / x_scale on demand
    (only the diagnostic cost getters of LinearSolver read it)."""

    def __init__(self, x_unscaled, x_scale):
        self._u, self._s = x_unscaled, x_scale

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._u / self._s, dtype=dtype)
