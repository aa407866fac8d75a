"""Linear operators of the reference API (``nsol.linear_operators``) backed by CUDA.

``LinearOperators{1,2,3}D`` keep the reference's constructor and method
signatures (nsol/linear_operators.py:18-247).  Every factory returns a pair of
callables ``(op, op_adj)`` exactly as the reference does, but each callable is
an ``Operator`` object that

* applied to a numpy array, runs the corresponding kernel of libnsol_b200 on
  the GPU (upload -> kernel -> download) and returns a fresh float64 array of
  the reference's output shape;
* applied to a ``_trace.Symbol`` during a solver's probing pass, describes
  itself, so that PrimalDualSolver / ADMMLinearSolver / TikhonovLinearSolver can
  keep the whole iteration on the device even when the application wrapped the
  operator in a lambda (nsol/application/run_denoising.py:104-107).

Semantics follow the reference: the gradient is a forward difference with a
ZERO boundary (``mode="constant"``, nsol/linear_operators.py:98-106) and its
exact adjoint; the blur is periodic (``mode="wrap"``, :60-68) and ``A_adj``
reuses the same mask (:63).  Other boundary modes are rejected.
"""
import ctypes as C
from abc import ABCMeta

import numpy as np

from nsol_b200 import _lib
from nsol_b200 import kernels as Kernels
from nsol_b200._device import as_double_p, run_array_op
from nsol_b200._trace import Symbol, is_symbol


class Operator(object):
    """Callable returned by the factories below; ``kind`` + parameters describe it."""

    kind = None

    def __init__(self, dimension, spacing):
        self.dimension = dimension
        self.spacing = np.atleast_1d(np.asarray(spacing, dtype=np.float64)).copy()

    def _check_ndim(self, x):
        if x.ndim != self.dimension:
            raise ValueError("%s operator of a %dD grid applied to a %dD array"
                             % (self.kind, self.dimension, x.ndim))


class GradientOperator(Operator):
    """grad(x) = concatenate((Dx x, Dy x[, Dz x])) on axis 0 -- nsol/linear_operators.py:121-144."""
    kind = "grad"

    def __call__(self, x):
        self._check_ndim(x)
        d = self.dimension
        out_shape = (d * x.shape[0],) + tuple(x.shape[1:])
        if is_symbol(x):
            return Symbol(("grad", x.expr, d, tuple(self.spacing), tuple(x.shape)), out_shape)
        grid = _lib.make_grid(x.shape, self.spacing)
        return run_array_op(
            lambda ctx, i, o, t: ctx.check(ctx.lib.nsol_grad(ctx.handle, C.byref(grid), i, o, None)),
            x, out_shape)


class GradientAdjointOperator(Operator):
    """grad_adj(p) = sum_k D_k^T p_k, blocks split on axis 0 -- nsol/linear_operators.py:158-169."""
    kind = "grad_adj"

    def __call__(self, p):
        self._check_ndim(p)
        d = self.dimension
        if p.shape[0] % d:
            raise ValueError("adjoint gradient: axis 0 (%d) is not a multiple of the dimension %d" % (p.shape[0], d))
        out_shape = (p.shape[0] // d,) + tuple(p.shape[1:])
        if is_symbol(p):
            return Symbol(("grad_adj", p.expr, d, tuple(self.spacing), out_shape), out_shape)
        grid = _lib.make_grid(out_shape, self.spacing)
        return run_array_op(
            lambda ctx, i, o, t: ctx.check(ctx.lib.nsol_grad_adj(ctx.handle, C.byref(grid), i, o, None)),
            p, out_shape)


class DifferenceOperator(Operator):
    """One D_k or D_k^T -- nsol/linear_operators.py:98-106, 193-201, 219-247."""

    def __init__(self, dimension, spacing, component, adjoint):
        Operator.__init__(self, dimension, spacing)
        self.component = component
        self.adjoint = adjoint
        self.kind = "diff_adj" if adjoint else "diff"

    def __call__(self, x):
        self._check_ndim(x)
        if is_symbol(x):
            return Symbol((self.kind, x.expr, self.dimension, tuple(self.spacing), tuple(x.shape), self.component), x.shape)
        grid = _lib.make_grid(x.shape, self.spacing)
        return run_array_op(
            lambda ctx, i, o, t: ctx.check(ctx.lib.nsol_diff(ctx.handle, C.byref(grid), self.component,
                                                             1 if self.adjoint else 0, i, o, None)),
            x, x.shape)


class ConvolutionOperator(Operator):
    """Periodic convolution with a fixed mask -- nsol/linear_operators.py:60-68.
    A separable mask (diagonal covariance) runs as d one-dimensional passes."""
    kind = "conv"

    def __init__(self, dimension, spacing, mask):
        Operator.__init__(self, dimension, spacing)
        self.mask = np.ascontiguousarray(mask, dtype=np.float64)
        if self.mask.ndim != dimension:
            raise ValueError("convolution mask must be %dD" % dimension)
        if any(s % 2 == 0 for s in self.mask.shape):
            raise ValueError("convolution mask extents must be odd")
        taps = Kernels.separable_taps(self.mask)
        self.taps = None if taps is None else [np.ascontiguousarray(t, dtype=np.float64) for t in taps]

    def __call__(self, x):
        self._check_ndim(x)
        if is_symbol(x):
            return Symbol(("conv", x.expr, self, tuple(x.shape)), x.shape)
        grid = _lib.make_grid(x.shape, self.spacing)
        if self.taps is not None:
            taps_arr = (_lib.c_double_p * 3)(*[as_double_p(t) for t in self.taps] + [None] * (3 - self.dimension))
            radius = (C.c_int32 * 3)(*[(t.size - 1) // 2 for t in self.taps] + [0] * (3 - self.dimension))
            return run_array_op(
                lambda ctx, i, o, t: ctx.check(ctx.lib.nsol_blur_sep(ctx.handle, C.byref(grid), taps_arr, radius, i, o, t, None)),
                x, x.shape, n_tmp=x.size)
        kshape = (C.c_int64 * 3)(*list(self.mask.shape) + [1] * (3 - self.dimension))
        return run_array_op(
            lambda ctx, i, o, t: ctx.check(ctx.lib.nsol_conv_wrap(ctx.handle, C.byref(grid), as_double_p(self.mask), kshape, i, o, None)),
            x, x.shape)


def _require_mode(mode, expected, what):
    if mode != expected:
        raise ValueError("%s: only mode=%r (the reference default) is implemented on the GPU, got %r"
                         % (what, expected, mode))


class LinearOperators(object):
    __metaclass__ = ABCMeta

    def __init__(self, dimension, spacing):
        self._dimension = dimension
        self._spacing = spacing
        self._kernels = getattr(Kernels, "Kernels%dD" % dimension)(spacing=spacing)

    def get_spacing(self):
        return self._spacing

    def get_dimension(self):
        return self._dimension

    def _spacing_vector(self):
        return self._kernels.get_spacing()

    def get_convolution_and_adjoint_convolution_operators(self, kernel, mode="wrap"):
        """(A, A_adj); the reference reuses the mask for the adjoint (nsol/linear_operators.py:63)."""
        _require_mode(mode, "wrap", "convolution")
        A = ConvolutionOperator(self._dimension, self._spacing_vector(), kernel)
        return A, A

    def get_gaussian_blurring_operators(self, cov, alpha_cut=3):
        """nsol/linear_operators.py:82-86."""
        kernel = self._kernels.get_gaussian(cov=cov, alpha_cut=alpha_cut)
        return self.get_convolution_and_adjoint_convolution_operators(kernel)

    def _difference_pair(self, component, mode):
        _require_mode(mode, "constant", "finite difference")
        if component >= self._dimension:
            raise AttributeError("d%s operators are not defined in %dD" % ("xyz"[component], self._dimension))
        s = self._spacing_vector()
        return (DifferenceOperator(self._dimension, s, component, False),
                DifferenceOperator(self._dimension, s, component, True))

    def get_dx_operators(self, mode="constant"):
        return self._difference_pair(0, mode)

    def get_gradient_operators(self, mode="constant"):
        """(grad, grad_adj) -- nsol/linear_operators.py:121-144."""
        _require_mode(mode, "constant", "gradient")
        s = self._spacing_vector()
        return GradientOperator(self._dimension, s), GradientAdjointOperator(self._dimension, s)


class LinearOperators1D(LinearOperators):
    def __init__(self, spacing=1):
        LinearOperators.__init__(self, dimension=1, spacing=spacing)


class LinearOperators2D(LinearOperators):
    def __init__(self, spacing=np.ones(2)):
        LinearOperators.__init__(self, dimension=2, spacing=spacing)

    def get_dy_operators(self, mode="constant"):
        return self._difference_pair(1, mode)


class LinearOperators3D(LinearOperators):
    def __init__(self, spacing=np.ones(3)):
        LinearOperators.__init__(self, dimension=3, spacing=spacing)

    def get_dy_operators(self, mode="constant"):
        return self._difference_pair(1, mode)

    def get_dz_operators(self, mode="constant"):
        return self._difference_pair(2, mode)
