"""Drive the UNMODIFIED reference (gift-surg/NSoL v0.1.14) -- TEST / BASELINE INFRASTRUCTURE.

Imports the reference package from oracle/_ref (staged by oracle/make_ref.py; present on the GPU
box) or, in the build container, from /root/reference, with the six-function ``pysitk`` stub of
oracle/pysitk_stub.  Solver objects are wired exactly as the reference's applications wire them
(nsol/application/run_denoising.py:95-154, run_deconvolution.py:104-152).  Used by bench.py's
``--impl reference`` arm / ``cpu_baseline`` leg, by tests/ and by the fixture generators; never by
nsol_b200/.
"""
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.path.join(HERE, "_ref"), os.environ.get("NSOL_REFERENCE", "/root/reference")]
_mods = {}


def reference_root():
    for root in _CANDIDATES:
        if os.path.isfile(os.path.join(root, "nsol", "primal_dual_solver.py")):
            return root
    return None


def available():
    return reference_root() is not None


def _import(name):
    """Import ``nsol.<name>`` of the reference (cached)."""
    if name in _mods:
        return _mods[name]
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is not staged: run `python oracle/make_ref.py` in the build container")
    stub = os.path.join(HERE, "pysitk_stub")
    for p in (root, stub):
        if p not in sys.path:
            sys.path.insert(0, p)
    mod = importlib.import_module("nsol." + name)
    if not os.path.abspath(mod.__file__).startswith(os.path.abspath(root)):
        raise RuntimeError("another `nsol` package shadows the reference: %s" % mod.__file__)
    _mods[name] = mod
    return mod


def module(name):
    return _import(name)


def linops(dim, spacing=None):
    lo = _import("linear_operators")
    cls = getattr(lo, "LinearOperators%dD" % dim)
    if spacing is None:
        return cls()
    return cls(spacing=np.asarray(spacing, dtype=float) if dim > 1 else float(np.atleast_1d(spacing)[0]))


def pd_solver(obs, reg="TV", data="L2", alpha=0.05, L2=8.0, iterations=10, alg_type="ALG2", spacing=None, x_scale=None):
    """PrimalDualSolver for denoising, wired as nsol/application/run_denoising.py:95-154."""
    pd = _import("primal_dual_solver")
    prox = _import("proximal_operators").ProximalOperators
    dim = obs.ndim
    b = obs.flatten()
    x0 = obs.flatten()
    x_scale = float(np.max(obs)) if x_scale is None else x_scale
    grad, grad_adj = linops(dim, spacing).get_gradient_operators()
    X_shape = obs.shape
    Z_shape = (dim * X_shape[0],) + tuple(X_shape[1:])
    D_1D = lambda x: grad(x.reshape(*X_shape)).flatten()
    D_adj_1D = lambda x: grad_adj(x.reshape(*Z_shape)).flatten()
    if data == "L1":
        prox_f = lambda x, tau: prox.prox_ell1_denoising(x, tau, x0=b, x_scale=x_scale)
    else:
        prox_f = lambda x, tau: prox.prox_ell2_denoising(x, tau, x0=b, x_scale=x_scale)
    prox_g_conj = {"TV": prox.prox_tv_conj, "HUBER": prox.prox_huber_conj, "TK1": lambda q, s: q / (1 + s)}[reg]
    return pd.PrimalDualSolver(prox_f=prox_f, prox_g_conj=prox_g_conj, B=D_1D, B_conj=D_adj_1D, L2=L2, x0=x0, alpha=alpha,
                               iterations=iterations, x_scale=x_scale, alg_type=alg_type)


def deconv_ops(shape, cov, spacing=None):
    """(A, A_adj, D, D_adj) as 1-D wrappers, nsol/application/run_deconvolution.py:109-129."""
    dim = len(shape)
    ops = linops(dim, spacing)
    A, A_adj = ops.get_gaussian_blurring_operators(cov)
    grad, grad_adj = ops.get_gradient_operators()
    Z_shape = (dim * shape[0],) + tuple(shape[1:])
    w = lambda op, sh: (lambda x: op(x.reshape(*sh)).flatten())
    return w(A, shape), w(A_adj, shape), w(grad, shape), w(grad_adj, Z_shape)


def admm_solver(obs, var, alpha=0.01, rho=0.1, iterations=50, iter_max=10, x_scale=None, spacing=None, b_reg=0):
    """ADMMLinearSolver wired as deconvolution_solver_parameter_study_interface.py:282-299."""
    admm = _import("admm_linear_solver")
    dim = obs.ndim
    cov = var if dim == 1 else np.diag(np.atleast_1d(var) * np.ones(dim))
    A, A_adj, D, D_adj = deconv_ops(obs.shape, cov, spacing)
    xs = float(np.max(obs)) if x_scale is None else x_scale
    return admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(), dimension=dim,
                                 b_reg=b_reg, alpha=alpha, rho=rho, iterations=iterations, iter_max=iter_max, x_scale=xs)


def timed_run(solver):
    """(seconds, get_x()) of solver.run() -- what BASELINE.md section 4 times."""
    t0 = time.perf_counter()
    solver.run()
    dt = time.perf_counter() - t0
    return dt, solver.get_x()
