"""Tikhonov-regularised linear least squares, solved by LSMR on the GPU.

API of ``nsol.tikhonov_linear_solver.TikhonovLinearSolver``
(nsol/tikhonov_linear_solver.py:25-280):

    min_x 1/2 ||A x - b||^2 + alpha/2 ||B x - b_reg||^2

solved as the stacked system [A; sqrt(alpha) B] x = [b; sqrt(alpha) b_reg] with
``scipy.sparse.linalg.lsmr(..., maxiter=iter_max, atol=0, btol=0)`` semantics (cold
start, result clipped to ``bounds``).  Only the reference default
``minimizer="lsmr"`` / ``data_loss="linear"`` exists here; the scipy.optimize
branches of the reference (:161-220) are out of scope and raise ``ValueError``.
"""
import numpy as np

from nsol_b200 import _lib
from nsol_b200.linear_solver import LinearSolver, acquire_lsmr_plan

EPS = 1e-10   # nsol/definitions.py:11


class TikhonovLinearSolver(LinearSolver):

    def __init__(self, A, A_adj, b, B, B_adj, x0, alpha=0.01, b_reg=0, data_loss="linear", data_loss_scale=1,
                 minimizer="lsmr", iter_max=10, x_scale=1, verbose=0, bounds=(0, np.inf), dtype=None):
        LinearSolver.__init__(self, A=A, A_adj=A_adj, b=b, x0=x0, alpha=alpha, iter_max=iter_max,
                              minimizer=minimizer, data_loss=data_loss, data_loss_scale=data_loss_scale,
                              x_scale=x_scale, verbose=verbose, dtype=dtype)
        self._B = B
        self._B_adj = B_adj
        self._b_reg = b_reg / self._x_scale
        self._bounds = bounds

    def get_B(self):
        return self._B

    def get_B_adj(self):
        return self._B_adj

    def get_b_reg(self):
        return self._b_reg * self._x_scale

    def _get_cost_regularization_term(self, x):
        return 0.5 * np.sum(self._B(x) ** 2)

    def _run(self):
        self._check_lsmr_only()
        if self._observer is not None:
            self._observer.add_x(self.get_x())
        info = self._probe_lsq(self._B, self._B_adj)
        n = self._x0.size
        if self._bounds is not None:
            self._x0 = np.clip(self._x0, self._bounds[0], self._bounds[1])   # :143 (x0 is not passed to lsmr)
        lo, hi = (-np.inf, np.inf) if self._bounds is None else (float(self._bounds[0]), float(self._bounds[1]))
        plan = acquire_lsmr_plan(self, info, self._dtype)       # kept across runs (parameter studies)
        ctx = plan.ctx
        b = np.ascontiguousarray(self._b, dtype=np.float64)
        rows = info["dim"] * n if info["b_kind"] == "grad" else n
        b_reg = None
        if np.ndim(self._b_reg) != 0 or float(self._b_reg) != 0.0:
            b_reg = np.ascontiguousarray(np.broadcast_to(np.asarray(self._b_reg, dtype=np.float64), (rows,)))
        x_out = ctx.result_empty(n, np.float64)
        ctx.check(ctx.lib.nsol_tikhonov_run_host(
            plan.handle, float(self._alpha), 1.0, float(self._x_scale), b.ctypes.data,
            b_reg.ctypes.data if b_reg is not None else None, int(self._iter_max), lo, hi,
            x_out.ctypes.data, None))
        self._set_result(x_out)
        if self._observer is not None:
            self._observer.add_x(self.get_x())
