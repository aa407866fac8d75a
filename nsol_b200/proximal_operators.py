"""Proximal maps of the reference API (``nsol.proximal_operators``) backed by CUDA.

Same static-method signatures as the reference (nsol/proximal_operators.py:43-159).
Called on numpy arrays they run ``nsol_prox_apply`` on the GPU; called on a
``_trace.Symbol`` (a solver's probing pass) they describe themselves so the
solver can fuse them into its iteration kernel.
"""
import numpy as np

from nsol_b200 import _lib
from nsol_b200._device import run_array_op
from nsol_b200._trace import Symbol, is_symbol


def _apply(kind, x, p0, p1=0.0, x0=None):
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    ctx = _lib.context()
    d_x0 = None
    if x0 is not None:
        x0 = np.ascontiguousarray(np.broadcast_to(np.asarray(x0, dtype=np.float64), x.shape))
        d_x0 = ctx.device_alloc(max(x0.nbytes, 8)).upload(x0)
    try:
        return run_array_op(
            lambda c, i, o, t: c.check(c.lib.nsol_prox_apply(c.handle, _lib.PROX[kind], _lib.F64, n, i,
                                                             d_x0.ptr if d_x0 else None, float(p0), float(p1), o, None)),
            x, x.shape)
    finally:
        if d_x0:
            d_x0.free()


class ProximalOperators(object):

    @staticmethod
    def prox_linear_least_squares(x, tau, A, A_adj, b, x0, iter_max=10, verbose=0, data_loss="linear",
                                  data_loss_scale=1, minimizer="lsmr", x_scale=1, bounds=(0, np.inf)):
        """argmin_y 1/2 ||A y - b||^2 + 1/(2 tau) ||y - x||^2 via a Tikhonov/LSMR solve with
        B = I, alpha = 1/tau, b_reg = x (nsol/proximal_operators.py:44-78)."""
        if is_symbol(x):
            return Symbol(("prox_lls", x.expr, float(tau), A, A_adj, b, x0, int(iter_max), data_loss,
                           minimizer, float(x_scale), bounds), x.shape)
        import nsol_b200.tikhonov_linear_solver as tk
        ident = lambda v: v.flatten()
        solver = tk.TikhonovLinearSolver(
            A=A, A_adj=A_adj, B=ident, B_adj=ident, x0=x0 / float(x_scale), b=b / float(x_scale), b_reg=x,
            alpha=1. / tau, iter_max=iter_max, verbose=verbose, x_scale=x_scale, data_loss=data_loss,
            data_loss_scale=data_loss_scale, minimizer=minimizer, bounds=bounds)
        solver.run()
        return solver.get_x()

    @staticmethod
    def prox_ell1_denoising(x, tau, x0, x_scale=1.):
        """nsol/proximal_operators.py:96-98."""
        if is_symbol(x):
            return Symbol(("prox_ell1", x.expr, float(tau), x0, float(x_scale)), x.shape)
        return _apply("ELL1", x, tau, x_scale, x0)

    @staticmethod
    def prox_ell2_denoising(x, tau, x0, x_scale=1.):
        """nsol/proximal_operators.py:118-120."""
        if is_symbol(x):
            return Symbol(("prox_ell2", x.expr, float(tau), x0, float(x_scale)), x.shape)
        return _apply("ELL2", x, tau, x_scale, x0)

    @staticmethod
    def prox_tv_conj(x, sigma):
        """Element-wise projection x / max(1, |x|) (nsol/proximal_operators.py:139-140)."""
        if is_symbol(x):
            return Symbol(("prox_tv_conj", x.expr, float(sigma)), x.shape)
        return _apply("TV_CONJ", x, sigma)

    @staticmethod
    def prox_huber_conj(x, sigma, gamma=0.05):
        """nsol/proximal_operators.py:157-159.  (The reference divides its argument in place;
        this returns a new array and leaves the argument untouched.)"""
        if is_symbol(x):
            return Symbol(("prox_huber_conj", x.expr, float(sigma), float(gamma)), x.shape)
        return _apply("HUBER_CONJ", x, sigma, gamma)
