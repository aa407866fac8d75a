#!/usr/bin/env python
"""Time the primal-dual deconvolution (prox_f = prox_linear_least_squares: one 10-iteration LSMR solve per PD iteration)
through the public API.    python tools/time_pd_deconv.py [--size 2048] [--iterations 10] [--iter-max 10] [--dtype float64]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--iter-max", type=int, default=10)
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import nsol_b200.linear_operators as lo
    import nsol_b200.primal_dual_solver as pd
    from nsol_b200.proximal_operators import ProximalOperators as prox
    n = args.size
    shape, zshape = (n, n), (2 * n, n)
    rng = np.random.RandomState(0)
    obs = rng.rand(n, n) * 200 + 20
    xs = float(obs.max())
    ops = lo.LinearOperators2D()
    A0, A0_adj = ops.get_gaussian_blurring_operators(np.eye(2))
    g, g_adj = ops.get_gradient_operators()
    A = lambda x: A0(x.reshape(*shape)).flatten()
    A_adj = lambda x: A0_adj(x.reshape(*shape)).flatten()
    D = lambda x: g(x.reshape(*shape)).flatten()
    D_adj = lambda x: g_adj(x.reshape(*zshape)).flatten()
    b = obs.flatten()
    s = pd.PrimalDualSolver(
        prox_f=lambda x, tau: prox.prox_linear_least_squares(x=x, tau=tau, A=A, A_adj=A_adj, b=b, x0=b, iter_max=args.iter_max, x_scale=xs),
        prox_g_conj=prox.prox_tv_conj, B=D, B_conj=D_adj, L2=8, x0=b, alpha=0.01, iterations=args.iterations, x_scale=xs, dtype=args.dtype)
    s.run()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        s.run()
        x = s.get_x()
    dt = (time.perf_counter() - t0) / args.reps
    print("pd-deconv %dx%d %s: %.2f ms per solve (%d PD x %d LSMR iterations), %.1f us per PD iteration, checksum %.6e"
          % (n, n, args.dtype, dt * 1e3, args.iterations, args.iter_max, dt * 1e6 / args.iterations, float(x[::997].sum())), flush=True)


if __name__ == "__main__":
    main()
