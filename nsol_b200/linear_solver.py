"""Shared state of the linear least-squares solvers (``nsol.linear_solver.LinearSolver``,
nsol/linear_solver.py:30-344): A/A_adj/b/alpha/data_loss/minimizer/iter_max, with b
stored as ``b / x_scale``.  Also hosts the probing that maps the user's A / B
callables onto the device-resident LSMR plan (include/nsol_b200.h: nsol_lsq_desc)."""
import ctypes as C
from abc import ABCMeta, abstractmethod

import numpy as np

from nsol_b200 import _lib
from nsol_b200 import _trace
from nsol_b200.solver import Solver

_LOSSES = ("linear", "soft_l1", "huber", "cauchy", "arctan")   # nsol/loss_functions.py:251-266


class LinearSolver(Solver):
    __metaclass__ = ABCMeta

    def __init__(self, A, A_adj, b, x0, alpha, x_scale, data_loss, data_loss_scale, minimizer, iter_max, verbose,
                 dtype=None):
        Solver.__init__(self, x0=x0, x_scale=x_scale, verbose=verbose)
        self._A = A
        self._A_adj = A_adj
        self._b = b / self._x_scale                       # nsol/linear_solver.py:83
        self._alpha = float(alpha)
        self._data_loss = data_loss
        self._data_loss_scale = float(data_loss_scale)
        self._minimizer = minimizer
        self._iter_max = iter_max
        self._dtype = dtype
        self._lsq = None

    def release(self):
        """Free the device memory held by this solver (the cached LSMR plan)."""
        release_lsmr_plan(self)

    def get_A(self):
        return self._A

    def get_A_adj(self):
        return self._A_adj

    def get_b(self):
        return np.array(self._b) * self._x_scale

    def set_alpha(self, alpha):
        self._alpha = alpha

    def get_alpha(self):
        return self._alpha

    def set_data_loss(self, data_loss):
        if data_loss not in _LOSSES:
            raise ValueError("data_loss must be in " + str(_LOSSES))
        self._data_loss = data_loss

    def get_data_loss(self):
        return self._data_loss

    def set_data_loss_scale(self, data_loss_scale):
        self._data_loss_scale = data_loss_scale

    def get_data_loss_scale(self):
        return self._data_loss_scale

    def set_minimizer(self, minimizer):
        self._minimizer = minimizer

    def get_minimizer(self):
        return self._minimizer

    def set_iter_max(self, iter_max):
        self._iter_max = iter_max

    def get_iter_max(self):
        return self._iter_max

    # -- diagnostics (evaluated with the GPU operator closures; linear loss only) ---------
    def get_total_cost(self):
        return self.get_cost_data_term() + self._alpha * self.get_cost_regularization_term()

    def get_cost_data_term(self):
        return self._get_cost_data_term(np.asarray(self._x))

    def get_ell2_cost_data_term(self):
        return self._get_cost_data_term(np.asarray(self._x))

    def get_cost_regularization_term(self):
        return self._get_cost_regularization_term(np.asarray(self._x))

    def _get_cost_data_term(self, x):
        if self._data_loss != "linear":
            raise ValueError("only the 'linear' data loss is implemented on the GPU")
        residual = self._A(x) - self._b
        return 0.5 * np.sum(residual ** 2)

    def print_statistics(self, fmt="%.3e"):
        print("Computational time: %s" % (self.get_computational_time()))
        print("Cost data term (f, loss=%s): " % self._data_loss + fmt % self.get_cost_data_term())
        print("Cost regularization term (g): " + fmt % self.get_cost_regularization_term())

    @abstractmethod
    def _get_cost_regularization_term(self, x):
        pass

    # -- probing -----------------------------------------------------------------------
    def _check_lsmr_only(self):
        # nsol/tikhonov_linear_solver.py:122-128
        if self._minimizer == "lsmr" and self._data_loss != "linear":
            raise ValueError("lsmr solver cannot be used with non-linear data loss")
        if self._minimizer == "lsq_linear" and self._data_loss != "linear":
            raise ValueError("lsq_linear solver cannot be used with non-linear data loss")
        if self._minimizer != "lsmr":
            raise ValueError("minimizer '%s' is not available in the CUDA backend: only 'lsmr' with the "
                             "'linear' data loss (the reference default) runs on the GPU" % self._minimizer)

    def _probe_lsq(self, B, B_adj):
        """Map (A, A_adj, B, B_adj) to an nsol_lsq_desc description (cached)."""
        if self._lsq is not None:
            return self._lsq
        n = self._x0.size
        info = probe_least_squares(self._A, self._A_adj, B, B_adj, n)
        self._lsq = info
        return info


def _classify_a(expr):
    if expr == ("arg",):
        return "identity", None, None
    if expr[0] == "conv" and expr[1] == ("arg",):
        return "conv", expr[2], expr[3]
    raise TypeError("unsupported data operator A; supported: " + _trace.SUPPORTED)


def probe_least_squares(A, A_adj, B, B_adj, n):
    a_kind, a_op, a_shape = _classify_a(_trace.probe(A, n).expr)
    adj_kind, adj_op, _ = _classify_a(_trace.probe(A_adj, n).expr)
    if a_kind != adj_kind or (a_kind == "conv" and not np.array_equal(a_op.mask, adj_op.mask)):
        raise TypeError("A_adj must be built from the same operator as A (the reference reuses the mask, "
                        "nsol/linear_operators.py:63)")
    shape, spacing, dim = None, None, None
    bexpr = _trace.probe(B, n).expr
    if bexpr == ("arg",):
        b_kind = "identity"
        rows = n
    elif bexpr[0] == "grad" and bexpr[1] == ("arg",):
        b_kind = "grad"
        _, _, dim, spacing, shape = bexpr
        rows = dim * n
    else:
        raise TypeError("unsupported regularisation operator B; supported: " + _trace.SUPPORTED)
    badj = _trace.probe(B_adj, rows).expr
    if b_kind == "identity" and badj != ("arg",):
        raise TypeError("B_adj must be the identity when B is")
    if b_kind == "grad" and not (badj[0] == "grad_adj" and badj[1] == ("arg",) and badj[2:5] == (dim, spacing, shape)):
        raise TypeError("B_adj must be the adjoint gradient of B's grid")
    if a_kind == "conv":
        if a_op.taps is None:
            raise TypeError("non-separable convolution masks are available as stand-alone operators only; "
                            "the LSMR path needs a separable (diagonal covariance) mask")
        if shape is None:
            shape, spacing, dim = tuple(a_shape), tuple(a_op.spacing), a_op.dimension
        elif tuple(a_shape) != tuple(shape):
            raise ValueError("A works on shape %s but B on shape %s" % (a_shape, shape))
    if shape is None:
        shape, spacing, dim = (n,), (1.0,), 1
    if int(np.prod(shape)) != n:
        raise ValueError("operators reshape x0 (%d values) to %s" % (n, (shape,)))
    return dict(a_kind=a_kind, a_op=a_op, b_kind=b_kind, shape=tuple(shape), spacing=tuple(spacing), dim=dim)


def acquire_lsmr_plan(owner, info, dtype):
    """The device plan of ``owner`` (a solver object) for this least-squares problem: created on first use and
    kept across run() calls -- a parameter study re-runs one solver object many times, and allocating /
    freeing ~14 device arrays costs several times the 512^2 ADMM solve itself.  A solver whose operators,
    grid or dtype changed gets a new plan.  Freed by ``release_lsmr_plan`` or with the solver object."""
    taps = None
    if info["a_kind"] == "conv":
        taps = tuple(tuple(float(v) for v in np.asarray(t).reshape(-1)) for t in info["a_op"].taps)
    key = (tuple(info["shape"]), tuple(info["spacing"]), info["a_kind"], info["b_kind"], taps, _lib.dtype_code(dtype))
    cached = getattr(owner, "_lsmr_plan_cache", None)
    if cached is not None and cached[0] == key and cached[1].handle is not None:
        return cached[1]
    release_lsmr_plan(owner)
    plan = LsmrPlan(info, dtype)
    owner._lsmr_plan_cache = (key, plan)
    return plan


def release_lsmr_plan(owner):
    cached = getattr(owner, "_lsmr_plan_cache", None)
    if cached is not None:
        cached[1].close()
        owner._lsmr_plan_cache = None


class LsmrPlan(object):
    """RAII wrapper of nsol_lsmr_plan."""

    def __init__(self, info, dtype):
        self.ctx = _lib.context()
        desc = _lib.LsqDesc()
        desc.grid = _lib.make_grid(info["shape"], info["spacing"], _lib.dtype_code(dtype), 1)
        desc.a_op = _lib.A_BLUR if info["a_kind"] == "conv" else _lib.A_IDENTITY
        desc.b_op = {"grad": _lib.B_GRAD, "identity": _lib.B_IDENTITY, "none": _lib.B_NONE}[info["b_kind"]]
        self._keep = []
        if info["a_kind"] == "conv":
            for a, t in enumerate(info["a_op"].taps):
                t = np.ascontiguousarray(t, dtype=np.float64)
                self._keep.append(t)
                desc.taps[a] = t.ctypes.data_as(_lib.c_double_p)
                desc.radius[a] = (t.size - 1) // 2
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.nsol_lsmr_plan_create(self.ctx.handle, C.byref(desc), C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle is not None:
            self.ctx.lib.nsol_lsmr_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
