#!/usr/bin/env python
"""Where the wall clock of ADMMLinearSolver.run() goes (host API, BASELINE config 3 shape): plan creation,
nsol_admm_run_host, plan destruction.    python tools/time_admm_host.py [--size 512]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    import nsol_b200.admm_linear_solver as admm
    import nsol_b200.linear_operators as lo
    import nsol_b200.linear_solver as ls
    from nsol_b200 import _lib
    n = args.size
    rng = np.random.RandomState(0)
    img = rng.rand(n, n) * 255
    ops = lo.LinearOperators2D()
    A, A_adj = ops.get_gaussian_blurring_operators(np.eye(2))
    grad, grad_adj = ops.get_gradient_operators()
    shape, zshape = img.shape, (2 * n, n)
    solver = admm.ADMMLinearSolver(
        A=lambda x: A(x.reshape(*shape)).flatten(), A_adj=lambda x: A_adj(x.reshape(*shape)).flatten(), b=img.flatten(),
        B=lambda x: grad(x.reshape(*shape)).flatten(), B_adj=lambda x: grad_adj(x.reshape(*zshape)).flatten(),
        x0=img.flatten(), dimension=2, alpha=0.01, rho=0.1, iterations=50, iter_max=10, x_scale=float(img.max()))
    spent = {}

    def wrap(obj, name, key):
        orig = getattr(obj, name)

        def timed(*a, **k):
            t = time.perf_counter()
            try:
                return orig(*a, **k)
            finally:
                spent[key] = spent.get(key, 0.0) + time.perf_counter() - t
        setattr(obj, name, timed)

    wrap(ls.LsmrPlan, "__init__", "plan_create")
    wrap(ls.LsmrPlan, "close", "plan_destroy")
    ctx = _lib.context()
    orig_run = ctx.lib.nsol_admm_run_host

    class Lib(object):
        def __getattr__(self, k):
            return getattr(ctx_lib, k)
    ctx_lib = ctx.lib

    def timed_run(*a):
        t = time.perf_counter()
        rc = orig_run(*a)
        spent["admm_run_host"] = spent.get("admm_run_host", 0.0) + time.perf_counter() - t
        return rc
    timed_run.restype = orig_run.restype
    ctx_lib.nsol_admm_run_host = timed_run
    for rep in range(args.reps):
        spent.clear()
        t0 = time.perf_counter()
        solver.run()
        x = solver.get_x()
        dt = time.perf_counter() - t0
        print("run %d: %.2f ms total | %s" % (rep, dt * 1e3, ", ".join("%s %.2f ms" % (k, v * 1e3) for k, v in spent.items())), flush=True)


if __name__ == "__main__":
    main()
