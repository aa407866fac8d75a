"""Primal-dual deconvolution: ``PrimalDualSolver`` whose ``prox_f`` is
``ProximalOperators.prox_linear_least_squares`` -- the reference's default TV-L2 / Huber-L2
deconvolution wiring (nsol/deconvolution_solver_parameter_study_interface.py:255-280,
303-325; exercised by tests/solvers_test.py:102-352).  Every primal-dual iteration contains a
cold-started LSMR solve on [A; sqrt(1/t) I]; the whole loop runs on the device
(``nsol_pd_deconv_run_host``)."""
import ctypes as C

import numpy as np

from nsol_b200 import _lib
from nsol_b200.linear_solver import acquire_lsmr_plan, probe_least_squares


def run_pd_deconvolution(solver, cfg):
    _, _, _, A, A_adj, b, x0, iter_max, data_loss, minimizer, prox_scale, bounds = cfg["lls"]
    if minimizer != "lsmr" or data_loss != "linear":
        raise ValueError("prox_linear_least_squares (CUDA): only minimizer='lsmr' with data_loss='linear' is implemented")
    if bounds is None or float(bounds[0]) != 0.0 or not np.isinf(bounds[1]):
        raise ValueError("prox_linear_least_squares (CUDA): only bounds=(0, inf) (the reference default) is implemented")
    n = solver._x0.size
    ident = lambda v: v.flatten()
    info = probe_least_squares(A, A_adj, ident, ident, n)
    if info["a_kind"] == "conv" and tuple(info["shape"]) != tuple(cfg["shape"]):
        raise ValueError("prox_linear_least_squares: A works on shape %s but B on %s" % (info["shape"], cfg["shape"]))
    info = dict(info, shape=tuple(cfg["shape"]), spacing=tuple(cfg["spacing"]), dim=cfg["dim"])
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64).reshape(-1))
    if b.size != n:
        raise ValueError("prox_linear_least_squares: b has %d values, x0 has %d" % (b.size, n))
    cfg_pd = dict(cfg, data="L2", b_scale=1.0)
    desc = solver._make_desc(cfg_pd, [float(solver._alpha)])
    x0s = np.ascontiguousarray(solver._x0, dtype=np.float64)
    iters = int(solver._iterations)
    x_out = _lib.context().result_empty(n, np.float64)
    its = np.empty((iters + 1, n), dtype=np.float64) if solver._observer is not None else None
    plan = acquire_lsmr_plan(solver, info, solver._dtype)       # kept across runs (parameter studies)
    ctx = plan.ctx
    ctx.check(ctx.lib.nsol_pd_deconv_run_host(
        plan.handle, C.byref(desc), iters, int(iter_max), float(prox_scale), b.ctypes.data, x0s.ctypes.data,
        x_out.ctypes.data, its.ctypes.data if its is not None else None, None))
    if its is not None:
        for i in range(iters + 1):
            solver._observer.add_x(np.array(its[i]))
    solver._set_result(x_out)
