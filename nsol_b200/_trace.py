"""Symbolic probing of user callables.

The reference solvers take arbitrary Python callables (``prox_f``,
``prox_g_conj``, ``B``, ``B_conj``, ``A``, ``A_adj``) and its applications wrap
the library closures in lambdas such as
``lambda x: grad(x.reshape(*X_shape)).flatten()``
(nsol/application/run_denoising.py:104-107).  To keep that API while running the
whole loop on the GPU, a solver calls each callable once with a ``Symbol`` -- a
stand-in for a device array that records which library operator it flowed
through.  Library closures (LinearOperators*, ProximalOperators) recognise a
``Symbol`` and return a new one describing themselves; ``reshape``/``flatten``
and scalar ``*``, ``/`` are tracked.  Anything else a callable does to a
``Symbol`` raises ``TypeError``: there is deliberately no CPU fallback.
"""
import numpy as np

SUPPORTED = ("LinearOperators{1,2,3}D gradient / adjoint gradient, Gaussian blur or convolution operators, "
             "identity (x.flatten()), ProximalOperators.prox_tv_conj / prox_huber_conj / "
             "prox_ell1_denoising / prox_ell2_denoising / prox_linear_least_squares, "
             "and q / (1 + sigma) (first-order Tikhonov dual prox)")


class UnsupportedCallable(TypeError):
    pass


class Symbol(object):
    """Stand-in for an array during probing.  ``expr`` is a nested tuple."""
    __array_priority__ = 1000.0

    def __init__(self, expr, shape):
        self.expr = expr
        self.shape = tuple(int(s) for s in shape)

    # -- array protocol subset the reference lambdas use
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = [int(s) for s in shape]
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1]))
            shape[shape.index(-1)] = self.size // max(known, 1)
        if int(np.prod(shape)) != self.size:
            raise ValueError("cannot reshape array of size %d into shape %s" % (self.size, tuple(shape)))
        return Symbol(self.expr, shape)

    def flatten(self):
        return Symbol(self.expr, (self.size,))

    ravel = flatten

    def copy(self):
        return self

    def __truediv__(self, other):
        return self._scale(other, True)

    def __mul__(self, other):
        return self._scale(other, False)

    __rmul__ = __mul__

    def _scale(self, other, divide):
        if isinstance(other, Symbol) or np.ndim(other) != 0:
            raise UnsupportedCallable("only scalar scaling of a traced array is supported; " + SUPPORTED)
        return Symbol(("scale", self.expr, float(other), divide), self.shape)

    def _unsupported(self, *a, **k):
        raise UnsupportedCallable(
            "this callable is not built from operators the CUDA backend knows; supported: " + SUPPORTED)

    __add__ = __radd__ = __sub__ = __rsub__ = __neg__ = __abs__ = __getitem__ = __iter__ = _unsupported
    __array__ = __len__ = __float__ = __rtruediv__ = __pow__ = _unsupported

    def __array_ufunc__(self, *a, **k):
        self._unsupported()

    def __array_function__(self, *a, **k):
        self._unsupported()


def is_symbol(x):
    return isinstance(x, Symbol)


def probe(fn, size, *extra):
    """Call ``fn(Symbol, *extra)`` and return the resulting Symbol."""
    try:
        out = fn(Symbol(("arg",), (int(size),)), *extra)
    except UnsupportedCallable:
        raise
    except (TypeError, AttributeError, ValueError) as e:
        raise UnsupportedCallable("callable could not be mapped to CUDA kernels (%s: %s); supported: %s"
                                  % (type(e).__name__, e, SUPPORTED))
    if not isinstance(out, Symbol):
        raise UnsupportedCallable("callable did not return an array derived from its argument; supported: " + SUPPORTED)
    return out


def strip_scale(expr):
    """Split nested ("scale", e, f, divide) wrappers: returns (inner expr, [(f, divide), ...])."""
    scales = []
    while expr[0] == "scale":
        scales.append((expr[2], expr[3]))
        expr = expr[1]
    return expr, scales
