"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the
API surface of the reference exists, probing of user callables, and error behaviour.
No compute calls (there is no GPU here)."""
import os
import re

import numpy as np
import pytest

import nsol_b200.linear_operators as lo
import nsol_b200.primal_dual_solver as pd
import nsol_b200.admm_linear_solver as admm
import nsol_b200.tikhonov_linear_solver as tk
from nsol_b200 import _lib, _trace, kernels
from nsol_b200.proximal_operators import ProximalOperators as prox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "nsol_b200.h")).read()
    declared = set(re.findall(r"\b(nsol_[a-z0-9_]+)\s*\(", header))
    declared -= {"nsol_diff_adj"}   # mentioned in prose only
    lib = _lib.load()
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.nsol_version() == 100


def test_struct_layouts_match_header():
    # field order of include/nsol_b200.h; sizes guard against silent padding drift
    import ctypes as C
    assert C.sizeof(_lib.Grid) == 4 + 4 + 24 + 24 + 4 + 4
    assert C.sizeof(_lib.PdDesc) == C.sizeof(_lib.Grid) + 16 + 5 * 8 + 8
    assert C.sizeof(_lib.LsqDesc) == C.sizeof(_lib.Grid) + 8 + 24 + 12 + 4


def _wrapped(shape, spacing=None):
    dim = len(shape)
    ops = getattr(lo, "LinearOperators%dD" % dim)() if spacing is None else \
        getattr(lo, "LinearOperators%dD" % dim)(spacing=np.asarray(spacing, float) if dim > 1 else spacing[0])
    grad, grad_adj = ops.get_gradient_operators()
    zshape = (dim * shape[0],) + tuple(shape[1:])
    D = lambda x: grad(x.reshape(*shape)).flatten()
    D_adj = lambda x: grad_adj(x.reshape(*zshape)).flatten()
    return ops, D, D_adj


def test_probe_denoising_wirings():
    shape = (5, 6, 7)
    obs = np.random.RandomState(0).rand(*shape)
    b = obs.flatten()
    xs = float(obs.max())
    _, D, D_adj = _wrapped(shape, [0.7, 1.3, 2.1])
    cases = [
        (lambda x, t: prox.prox_ell2_denoising(x, t, x0=b, x_scale=xs), prox.prox_tv_conj, "TV", "L2"),
        (lambda x, t: prox.prox_ell1_denoising(x, t, x0=b, x_scale=xs), prox.prox_huber_conj, "HUBER", "L1"),
        (lambda x, t: prox.prox_ell2_denoising(x, t, x0=b, x_scale=xs), lambda q, s: q / (1 + s), "TK1", "L2"),
    ]
    for prox_f, prox_g, reg, data in cases:
        s = pd.PrimalDualSolver(prox_f=prox_f, prox_g_conj=prox_g, B=D, B_conj=D_adj, L2=8, x0=b, alpha=0.05, x_scale=xs)
        cfg = s._probe()
        assert (cfg["reg"], cfg["data"], cfg["dim"], cfg["shape"]) == (reg, data, 3, shape)
        assert cfg["spacing"] == (0.7, 1.3, 2.1) and cfg["b_scale"] == xs
        assert cfg["b"] is not None and cfg["b"].size == b.size


def test_unsupported_callables_raise_typeerror():
    shape = (8, 9)
    b = np.ones(72)
    _, D, D_adj = _wrapped(shape)
    good_f = lambda x, t: prox.prox_ell2_denoising(x, t, x0=b)
    for kw in (dict(prox_f=lambda x, t: x + 1.0), dict(prox_g_conj=lambda q, s: np.tanh(q)),
               dict(B=lambda x: x.flatten()), dict(B_conj=lambda x: x[:72])):
        args = dict(prox_f=good_f, prox_g_conj=prox.prox_tv_conj, B=D, B_conj=D_adj, L2=8, x0=b)
        args.update(kw)
        with pytest.raises(TypeError):
            pd.PrimalDualSolver(**args)._probe()


def test_solver_api_surface_and_errors():
    shape = (8, 9)
    b = np.arange(72.0)
    _, D, D_adj = _wrapped(shape)
    s = pd.PrimalDualSolver(prox_f=lambda x, t: prox.prox_ell2_denoising(x, t, x0=b), prox_g_conj=prox.prox_tv_conj,
                            B=D, B_conj=D_adj, L2=8, x0=b, x_scale=4.0)
    for name in ("alpha", "L2", "alg_type", "iterations", "x_scale", "verbose"):
        assert hasattr(s, "set_" + name) and hasattr(s, "get_" + name)
    assert np.array_equal(s.get_x0(), (b / 4.0) * 4.0) and np.array_equal(s.get_x(), s.get_x0())
    s.set_alpha(0.3)
    assert s.get_alpha() == 0.3 and s.get_computational_time().total_seconds() == 0
    # nsol/solver.py:149-150: non 1-D x0 -> ValueError before anything touches the GPU
    s2 = pd.PrimalDualSolver(prox_f=None, prox_g_conj=None, B=None, B_conj=None, L2=8, x0=b.reshape(shape))
    with pytest.raises(ValueError):
        s2.run()
    # no CUDA device here: the product path must fail loudly, not fall back
    with pytest.raises(RuntimeError):
        s.run()


def test_linear_solver_probe_and_errors():
    shape = (16, 12)
    n = 16 * 12
    ops, D, D_adj = _wrapped(shape)
    A, A_adj = ops.get_gaussian_blurring_operators(np.eye(2))
    A1 = lambda x: A(x.reshape(*shape)).flatten()
    A1_adj = lambda x: A_adj(x.reshape(*shape)).flatten()
    b = np.ones(n)
    s = admm.ADMMLinearSolver(A=A1, A_adj=A1_adj, b=b, B=D, B_adj=D_adj, x0=b, dimension=2, x_scale=2.0)
    info = s._probe_lsq(s._B, s._B_adj)
    assert info["a_kind"] == "conv" and info["b_kind"] == "grad" and info["shape"] == shape
    assert [t.size for t in info["a_op"].taps] == [7, 7]
    ident = lambda x: x.flatten()
    t = tk.TikhonovLinearSolver(A=A1, A_adj=A1_adj, b=b, B=ident, B_adj=ident, x0=b)
    assert t._probe_lsq(t._B, t._B_adj)["b_kind"] == "identity"
    # nsol/tikhonov_linear_solver.py:122-128
    t2 = tk.TikhonovLinearSolver(A=A1, A_adj=A1_adj, b=b, B=ident, B_adj=ident, x0=b, data_loss="huber")
    with pytest.raises(ValueError):
        t2.run()
    assert np.array_equal(s.get_b(), (b / 2.0) * 2.0)


def test_operator_argument_errors():
    with pytest.raises(ValueError):
        kernels.Kernels2D(spacing=np.ones(3))                    # nsol/kernels.py:22-23
    with pytest.raises(ValueError):
        kernels.Kernels2D().get_gaussian(np.eye(3))              # nsol/kernels.py:122-124
    with pytest.raises(ValueError):
        lo.LinearOperators2D().get_gradient_operators(mode="wrap")
    grad, _ = lo.LinearOperators2D().get_gradient_operators()
    with pytest.raises(ValueError):
        grad(_trace.Symbol(("arg",), (4, 5, 6)))


def test_gaussian_mask_is_separable_for_diagonal_cov():
    m = kernels.Kernels3D(spacing=np.array([1.0, 1.5, 0.7])).get_gaussian(np.diag([0.8, 1.7, 1.2]))
    taps = kernels.separable_taps(m)
    rec = np.multiply.outer(np.multiply.outer(taps[0], taps[1]), taps[2])
    assert np.max(np.abs(rec - m)) < 1e-16
    assert kernels.separable_taps(kernels.Kernels2D().get_gaussian(np.array([[1.0, 0.6], [0.6, 2.0]]))) is None


def test_lsmr_plan_cache_keys_on_operators_grid_and_dtype(monkeypatch):
    """linear_solver.acquire_lsmr_plan: a solver object keeps its device plan across run() calls and gets a new one
    (the old one released) when the blur taps, the grid, B or the dtype change."""
    import nsol_b200.linear_solver as ls
    made, closed = [], []

    class FakePlan(object):
        def __init__(self, info, dtype):
            self.handle = object()
            made.append((info["shape"], dtype))

        def close(self):
            closed.append(self)
            self.handle = None

    monkeypatch.setattr(ls, "LsmrPlan", FakePlan)

    class Op(object):
        def __init__(self, taps):
            self.taps = taps

    class Owner(object):
        pass

    t3 = [np.array([0.25, 0.5, 0.25])] * 2
    info = dict(shape=(8, 6), spacing=(1.0, 1.0), a_kind="conv", b_kind="grad", a_op=Op(t3))
    o = Owner()
    p1 = ls.acquire_lsmr_plan(o, info, None)
    assert ls.acquire_lsmr_plan(o, dict(info, a_op=Op([t.copy() for t in t3])), "float64") is p1     # equal taps, same dtype
    p2 = ls.acquire_lsmr_plan(o, dict(info, a_op=Op([np.array([0.2, 0.6, 0.2])] * 2)), None)
    assert p2 is not p1 and closed == [p1]
    p3 = ls.acquire_lsmr_plan(o, dict(info, a_op=Op([np.array([0.2, 0.6, 0.2])] * 2)), "float32")
    p4 = ls.acquire_lsmr_plan(o, dict(info, a_op=Op([np.array([0.2, 0.6, 0.2])] * 2), shape=(8, 8)), "float32")
    p5 = ls.acquire_lsmr_plan(o, dict(info, a_op=Op([np.array([0.2, 0.6, 0.2])] * 2), shape=(8, 8), b_kind="identity"), "float32")
    assert len({id(p) for p in (p1, p2, p3, p4, p5)}) == 5 and closed == [p1, p2, p3, p4]
    ls.release_lsmr_plan(o)
    assert closed[-1] is p5 and o._lsmr_plan_cache is None
    ls.release_lsmr_plan(o)          # idempotent
    assert len(made) == 5
