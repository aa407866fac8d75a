#!/usr/bin/env python
"""How close can ANY float32 implementation of the LSMR-based solvers get to the float64 reference?

The float32 mode of nsol_b200 stores u, v, h, hbar, x (and the ADMM split variables) in float32 and runs
the scalar recurrences in float64.  This experiment runs the oracle's restatement of the reference
algorithms on the CPU with every VECTOR rounded to float32 after each operation (norms and scalars in
float64 -- the most favourable float32 arrangement, the one the CUDA kernels use) and reports the
relative max-abs deviation from the float64 run, next to PSNR / SSIM / NCC of both against the clean
image.  It shows whether north_star's 1e-4 is reachable for ADMM / PD-deconvolution at all, or whether
the deviation is inherent to float32 storage of the Krylov vectors (the solves are cold-started, clipped
and fed back 50 times).  Output committed under profiles/r2_fp32_bound.md.

    python tools/fp32_bound_experiment.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nsol_oracle as orc  # noqa: E402


class F32Ops(object):
    """Operators whose outputs are rounded to float32 (inputs arrive as float32-representable float64)."""

    def __init__(self, ops):
        self.ops = ops

    def wrap(self, f):
        return lambda x: f(x).astype(np.float32).astype(np.float64)


def lsmr_f32(matvec, rmatvec, b, n, maxiter):
    """scipy's LSMR recurrences (oracle.lsmr) with every vector update rounded to float32."""
    r32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    u = r32(b)
    normb = np.linalg.norm(u)
    x = np.zeros(n)
    beta = normb
    if beta > 0:
        u = r32(u * (1 / beta))
        v = r32(rmatvec(u))
        alpha = np.linalg.norm(v)
    else:
        v = np.zeros(n)
        alpha = 0
    if alpha > 0:
        v = r32(v * (1 / alpha))
    itn = 0
    zetabar = alpha * beta
    alphabar = alpha
    rho = rhobar = cbar = 1
    sbar = 0
    h = v.copy()
    hbar = np.zeros(n)
    if alpha * beta == 0:
        return x
    while itn < maxiter:
        itn += 1
        u = r32(u * -alpha)
        u = r32(u + matvec(v))
        beta = np.linalg.norm(u)
        if beta > 0:
            u = r32(u * (1 / beta))
            v = r32(v * -beta)
            v = r32(v + rmatvec(u))
            alpha = np.linalg.norm(v)
            if alpha > 0:
                v = r32(v * (1 / alpha))
        chat, shat, alphahat = orc.sym_ortho(alphabar, 0.0)
        rhoold = rho
        c, s, rho = orc.sym_ortho(alphahat, beta)
        thetanew = s * alpha
        alphabar = c * alpha
        rhobarold = rhobar
        thetabar = sbar * rho
        cbar, sbar, rhobar = orc.sym_ortho(cbar * rho, thetanew)
        zeta = cbar * zetabar
        zetabar = -sbar * zetabar
        hbar = r32(hbar * (-(thetabar * rho / (rhoold * rhobarold))))
        hbar = r32(hbar + h)
        x = r32(x + (zeta / (rho * rhobar)) * hbar)
        h = r32(h * (-(thetanew / rho)))
        h = r32(h + v)
    return x


def admm_f32(A, A_adj, B, B_adj, b, x0, dim, alpha, rho, iterations, iter_max, x_scale):
    r32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    x = r32(np.asarray(x0, dtype=np.float64) / x_scale)
    bs = r32(np.asarray(b, dtype=np.float64) / x_scale)
    v = r32(B(x))
    w = np.zeros_like(v)
    sa = np.sqrt(rho)
    n = x.size
    for _ in range(iterations):
        rhs = np.concatenate((bs, r32(sa * r32(v - w))))
        fw = lambda y: np.concatenate((r32(A(y)), r32(sa * r32(B(y)))))
        bw = lambda y: r32(r32(A_adj(y[:n])) + r32(sa * r32(B_adj(y[n:]))))
        x = r32(np.clip(lsmr_f32(fw, bw, rhs, n, iter_max), 0, np.inf))
        t = r32(r32(B(x)) + w)
        v = r32(orc.admm_shrink_iso(t, alpha / rho, dim))
        w = r32(t - v)
    return x * x_scale


def pd_f32(obs, alpha, L2, iters, x_scale):
    """The primal-dual loop (nsol/primal_dual_solver.py:232-261, TV-L2 wiring) with every array and operation in
    numpy float32; step sizes from the float64 schedule, like the CUDA kernels."""
    f = np.float32
    shape, dim = obs.shape, obs.ndim
    b = (obs.reshape(-1) / x_scale).astype(f)
    x, xm, p = b.copy(), b.copy(), np.zeros(dim * b.size, dtype=f)
    zshape = (dim * shape[0],) + shape[1:]
    for sigma, tau, tl, theta in orc.pd_schedule("ALG2", L2, alpha, iters):
        g = orc.grad(xm.reshape(shape)).reshape(-1).astype(f)
        q = p + f(sigma) * g
        p = (q / np.maximum(f(1), np.abs(q))).astype(f)
        d = orc.grad_adj(p.reshape(zshape), None, dim).reshape(-1).astype(f)
        y = x - f(tau) * d
        xn = ((y + f(tl) * b) / (f(1) + f(tl))).astype(f)
        xm = (xn + f(theta) * (xn - x)).astype(f)
        x = xn
    return x.astype(np.float64) * x_scale


def measures(x, clean):
    return orc.psnr(x, clean), orc.ssim_1d(x, clean), orc.ncc(x, clean)


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "inputs.npz"))
    print("| case | rel. max-abs deviation float32 vs float64 (CPU, same algorithm) | PSNR f64 / f32 | SSIM f64 / f32 | NCC f64 / f32 | 3 decimals equal |")
    print("|---|---|---|---|---|---|")
    for name, clean, outer in (("ADMM TV-L2, lena 128x128 crop, 50 x 10 (config 3 parameters)", z["lena_512"][192:320, 192:320].astype(np.float64), 50),
                               ("ADMM TV-L2, lena 256x256 crop, 50 x 10", z["lena_512"][128:384, 128:384].astype(np.float64), 50),
                               ("ADMM TV-L2, lena 128x128 crop, 10 x 10", z["lena_512"][192:320, 192:320].astype(np.float64), 10)):
        shape = clean.shape
        A, A_adj, D, D_adj = orc.deconvolution_operators(shape, np.eye(2))
        obs = orc.add_gaussian_noise(A(clean.reshape(-1)).reshape(shape), 0.05, seed=1)
        xs = float(obs.max())
        x64 = orc.admm_tv(A, A_adj, D, D_adj, obs.reshape(-1), obs.reshape(-1), 2, alpha=0.01, rho=0.1, iterations=outer, iter_max=10, x_scale=xs)
        x32 = admm_f32(A, A_adj, D, D_adj, obs.reshape(-1), obs.reshape(-1), 2, 0.01, 0.1, outer, 10, xs)
        dev = np.max(np.abs(x32 - x64)) / np.max(np.abs(x64))
        m64, m32 = measures(x64, clean.reshape(-1)), measures(x32, clean.reshape(-1))
        same = all(round(a, 3) == round(b, 3) for a, b in zip(m64, m32))
        print("| %s | %.2e | %.4f / %.4f | %.5f / %.5f | %.6f / %.6f | %s |" % (name, dev, m64[0], m32[0], m64[1], m32[1], m64[2], m32[2], same))


def main_pd():
    z = np.load(os.path.join(ROOT, "tests", "golden", "inputs.npz"))
    ph = z["shepp_logan_64"].astype(np.float64)
    print()
    print("| primal-dual TV-L2 (config 4 wiring: alpha 0.05, L2 8, ALG2) | rel. max-abs deviation float32 vs float64 (CPU) | PSNR f64 / f32 | SSIM f64 / f32 | NCC f64 / f32 | 3 decimals equal |")
    print("|---|---|---|---|---|---|")
    for rep, iters in ((1, 30), (1, 100), (2, 30), (2, 100)):
        vol = ph
        for ax in range(3):
            vol = np.repeat(vol, rep, ax)
        voln = orc.add_gaussian_noise(vol, 0.05, seed=1)
        xs = float(voln.max())
        x64 = orc.primal_dual_denoise(voln.reshape(-1), voln.shape, reg="TV", data="L2", alpha=0.05, L2=8, iterations=iters, x_scale=xs)
        x32 = pd_f32(voln, 0.05, 8.0, iters, xs)
        dev = np.max(np.abs(x32 - x64)) / np.max(np.abs(x64))
        m64, m32 = measures(x64, vol.reshape(-1)), measures(x32, vol.reshape(-1))
        same = all(round(a, 3) == round(b, 3) for a, b in zip(m64, m32))
        print("| %d^3, %d iterations | %.2e | %.4f / %.4f | %.5f / %.5f | %.6f / %.6f | %s |" % (64 * rep, iters, dev, m64[0], m32[0], m64[1], m32[1], m64[2], m32[2], same), flush=True)


if __name__ == "__main__":
    main_pd()
    main()
