"""Regulariser cost values (``nsol.prior_measures.PriorMeasures``, nsol/prior_measures.py:16-52),
evaluated on the GPU by one fused reduction (``nsol_prior_stats``).  ``D`` must be a
``LinearOperators*`` gradient operator (possibly wrapped in the usual reshape/flatten lambda)."""
import ctypes as C

import numpy as np

from nsol_b200 import _lib
from nsol_b200 import _trace


def _stats(x, D, gamma=0.05):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if D is None:
        shape, spacing = (x.size,), None
    else:
        expr = _trace.probe(D, x.size).expr if x.ndim == 1 else D(_trace.Symbol(("arg",), x.shape)).expr
        if expr[0] != "grad" or expr[1] != ("arg",):
            raise TypeError("prior measures need a LinearOperators gradient operator; supported: " + _trace.SUPPORTED)
        shape, spacing = expr[4], expr[3]
    ctx = _lib.context()
    grid = _lib.make_grid(shape, spacing, _lib.F64, 1)
    dx = ctx.device_alloc(max(x.nbytes, 8)).upload(x)
    out = np.empty(4, dtype=np.float64)
    try:
        ctx.check(ctx.lib.nsol_prior_stats(ctx.handle, C.byref(grid), dx.ptr, 1.0, float(gamma), out.ctypes.data, None))
    finally:
        dx.free()
    return out


class PriorMeasures(object):

    @staticmethod
    def zeroth_order_tikhonov(x):
        return 0.5 * _stats(x, None)[3]

    @staticmethod
    def first_order_tikhonov(x, D):
        return 0.5 * _stats(x, D)[1]

    @staticmethod
    def total_variation(x, D, dimension):
        return _stats(x, D)[0]

    @staticmethod
    def huber(x, D, dimension, gamma=0.05):
        return _stats(x, D, gamma)[2] / (2. * gamma)
