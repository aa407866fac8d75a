#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python oracle/gen_golden.py

The reference (gift-surg/NSoL v0.1.14) is pure Python; it is imported from
/root/reference with the 6-function ``pysitk`` stub in oracle/pysitk_stub.  The
solver objects are wired exactly as nsol/application/run_denoising.py:95-154
and run_deconvolution.py:104-152 wire them.  Nothing from the reference is
copied: only its *outputs* (and uint8 copies of its bundled test images, which
are the inputs those outputs belong to) are stored.

Fixtures written (all float64 unless noted):
  inputs.npz   uint8 test images of the reference's data/ directory
  ops.npz      grad / grad_adj / blur / gaussian-mask outputs
  pd.npz       PrimalDualSolver results (TV|Huber|TK1 x L1|L2 x ALG2|ALG3|AHMOD, 1D/2D/3D)
  lsmr.npz     TikhonovLinearSolver (lsmr) and ADMMLinearSolver results
  manifest.json  parameters of every case
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NSOL_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "pysitk_stub"))
sys.path.insert(1, REF)

import nsol.linear_operators as lo  # noqa: E402
import nsol.kernels as kern  # noqa: E402
import nsol.primal_dual_solver as pd  # noqa: E402
import nsol.admm_linear_solver as admm  # noqa: E402
import nsol.tikhonov_linear_solver as tk  # noqa: E402
import nsol.noise as noise  # noqa: E402
from nsol.proximal_operators import ProximalOperators as prox  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def read_png(name):
    from PIL import Image
    im = Image.open(os.path.join(REF, "data", name))
    assert im.mode == "L", (name, im.mode)
    return np.array(im)


def read_phantom():
    raw = gzip.open(os.path.join(REF, "data", "3D_SheppLoganPhantom_64.nii.gz")).read()
    a = np.frombuffer(raw[352:], dtype=np.float64).reshape(64, 64, 64)
    assert np.array_equal(a, a.astype(np.uint8))
    return a.astype(np.uint8)


def linops(dim, spacing=None):
    cls = getattr(lo, "LinearOperators%dD" % dim)
    if spacing is None:
        return cls()
    return cls(spacing=np.asarray(spacing, dtype=float) if dim > 1 else float(spacing[0]))


def wrap_1d(op, shape_in):
    return lambda x: op(x.reshape(*shape_in)).flatten()


def ref_pd(obs, reg, data, alpha, L2, iterations, alg_type="ALG2", spacing=None,
           x_scale=None, iterates=False):
    """run_denoising.py:95-154 + :187-188."""
    dim = obs.ndim
    b = obs.flatten()
    x0 = obs.flatten()
    x_scale = float(np.max(obs)) if x_scale is None else x_scale
    grad, grad_adj = linops(dim, spacing).get_gradient_operators()
    X_shape = obs.shape
    Z_shape = grad(obs).shape
    D_1D = lambda x: grad(x.reshape(*X_shape)).flatten()
    D_adj_1D = lambda x: grad_adj(x.reshape(*Z_shape)).flatten()
    if data == "L1":
        prox_f = lambda x, tau: prox.prox_ell1_denoising(x, tau, x0=b, x_scale=x_scale)
    else:
        prox_f = lambda x, tau: prox.prox_ell2_denoising(x, tau, x0=b, x_scale=x_scale)
    prox_g_conj = {"TV": prox.prox_tv_conj, "HUBER": prox.prox_huber_conj,
                   "TK1": lambda q, s: q / (1 + s)}[reg]
    solver = pd.PrimalDualSolver(prox_f=prox_f, prox_g_conj=prox_g_conj, B=D_1D, B_conj=D_adj_1D,
                                 L2=L2, x0=x0, alpha=alpha, iterations=iterations,
                                 x_scale=x_scale, alg_type=alg_type)
    xs = []
    if iterates:
        class Obs(object):
            def add_x(self, x):
                xs.append(x)

            def set_computational_time(self, t):
                pass
        solver.set_observer(Obs())
    solver.run()
    if iterates:
        return solver.get_x(), np.array(xs)
    return solver.get_x()


def ref_deconv_ops(shape, cov, spacing=None):
    """run_deconvolution.py:109-129."""
    dim = len(shape)
    ops = linops(dim, spacing)
    A, A_adj = ops.get_gaussian_blurring_operators(cov)
    grad, grad_adj = ops.get_gradient_operators()
    Z_shape = grad(np.zeros(shape)).shape
    return (wrap_1d(A, shape), wrap_1d(A_adj, shape), wrap_1d(grad, shape),
            wrap_1d(grad_adj, Z_shape))


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}

    # ------------------------------------------------------------ inputs
    lena_noise = read_png("2D_Lena_256_noise.png")
    brainweb = read_png("2D_BrainWeb.png")
    lena512 = read_png("2D_Lena_512.png")
    man1024 = read_png("2D_Man_1024.png")
    phantom = read_phantom()
    np.savez_compressed(os.path.join(OUT, "inputs.npz"), lena_256_noise=lena_noise,
                        brainweb=brainweb, lena_512=lena512, man_1024=man1024,
                        shepp_logan_64=phantom)

    # ------------------------------------------------------------ operators
    ops_out = {}
    ops_manifest = {}
    rng = np.random.RandomState(1234)
    op_cases = [
        ("g1", (50,), None), ("g1s", (37,), [0.7]),
        ("g2", (50, 50), None), ("g2s", (23, 31), [0.7, 1.3]),
        ("g3", (10, 50, 50), None), ("g3s", (7, 9, 11), [0.7, 1.3, 2.1]),
    ]
    for name, shape, spacing in op_cases:
        dim = len(shape)
        x = rng.rand(*shape)
        grad, grad_adj = linops(dim, spacing).get_gradient_operators()
        g = grad(x)
        y = rng.rand(*g.shape)
        ops_out[name + "/x"] = x
        ops_out[name + "/grad"] = g
        ops_out[name + "/y"] = y
        ops_out[name + "/grad_adj"] = grad_adj(np.array(y))
        ops_manifest[name] = {"shape": shape, "spacing": spacing}
    blur_cases = [
        ("b1", (50,), 1.5, None), ("b1s", (41,), 2.0, [0.8]),
        ("b2", (50, 50), [1.5, 1.5], None), ("b2a", (23, 31), [1.0, 2.5], [0.9, 1.4]),
        ("b3", (10, 24, 20), [1.0, 1.0, 1.0], None), ("b3a", (9, 14, 12), [0.8, 1.7, 1.2], [1.0, 1.5, 0.7]),
        ("b2small", (5, 4), [1.0, 1.0], None),
    ]
    for name, shape, var, spacing in blur_cases:
        dim = len(shape)
        x = rng.rand(*shape)
        cov = var if dim == 1 else np.diag(var)
        ops = linops(dim, spacing)
        A, A_adj = ops.get_gaussian_blurring_operators(cov)
        if dim == 1:
            kernel = kern.Kernels1D(spacing=spacing[0] if spacing else 1).get_gaussian(cov)
        else:
            kcls = getattr(kern, "Kernels%dD" % dim)
            kernel = (kcls(spacing=np.asarray(spacing, float)) if spacing else kcls()).get_gaussian(cov)
        ops_out[name + "/x"] = x
        ops_out[name + "/kernel"] = kernel
        ops_out[name + "/A"] = A(x)
        ops_out[name + "/A_adj"] = A_adj(x)
        ops_manifest[name] = {"shape": shape, "var": var, "spacing": spacing}
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **ops_out)
    manifest["ops"] = ops_manifest

    # ------------------------------------------------------------ primal-dual
    pd_out = {}
    pd_manifest = {}
    spike = np.ones(50) * 50
    spike[5], spike[16], spike[23], spike[30] = 10, 100, 150, 20  # tests/solvers_test.py:78-82
    spike_noisy = noise.Noise(spike, seed=1)
    spike_noisy.add_gaussian_noise(noise_level=0.05)
    spike_noisy = spike_noisy.get_noisy_data()
    bw_crop = brainweb[60:97, 50:79].astype(np.float64)          # 37 x 29, odd sizes
    ph3 = phantom[20:41, 16:35, 10:33].astype(np.float64)        # 21 x 19 x 23
    ph3n = noise.Noise(ph3, seed=1)
    ph3n.add_gaussian_noise(noise_level=0.05)
    ph3n = ph3n.get_noisy_data()
    ph32 = phantom[::2, ::2, ::2].astype(np.float64)             # 32^3
    ph32n = noise.Noise(ph32, seed=1)
    ph32n.add_gaussian_noise(noise_level=0.05)
    ph32n = ph32n.get_noisy_data()
    man_sp = noise.Noise(man1024[256:384, 512:640].astype(np.float64), seed=1)
    man_sp.add_salt_and_pepper_noise(salt_vs_pepper=0.5, amount=0.1)
    man_sp = man_sp.get_noisy_data()                             # 128 x 128 s&p

    pd_inputs = {"spike1d": spike_noisy, "bw2d": bw_crop, "ph3d": ph3n, "ph32": ph32n,
                 "man_sp": man_sp, "lena": lena_noise.astype(np.float64)}
    for k, v in pd_inputs.items():
        pd_out["in/" + k] = v

    def add_pd(name, inp, **kw):
        x = ref_pd(pd_inputs[inp], **kw)
        pd_out[name] = x
        pd_manifest[name] = dict(input=inp, **kw)

    for reg in ("TV", "HUBER", "TK1"):
        for data in ("L1", "L2"):
            alpha = 0.6 if data == "L1" else 0.05
            add_pd("1d_%s_%s" % (reg, data), "spike1d", reg=reg, data=data, alpha=alpha, L2=4, iterations=30)
            add_pd("2d_%s_%s" % (reg, data), "bw2d", reg=reg, data=data, alpha=alpha, L2=8, iterations=30)
            add_pd("3d_%s_%s" % (reg, data), "ph3d", reg=reg, data=data, alpha=alpha, L2=8, iterations=20)
    for alg in ("ALG3", "ALG2_AHMOD"):
        add_pd("2d_TV_L2_%s" % alg, "bw2d", reg="TV", data="L2", alpha=0.05, L2=8, iterations=30, alg_type=alg)
        add_pd("3d_HUBER_L1_%s" % alg, "ph3d", reg="HUBER", data="L1", alpha=0.6, L2=12, iterations=20, alg_type=alg)
    add_pd("2d_TV_L2_spacing", "bw2d", reg="TV", data="L2", alpha=0.05, L2=8, iterations=25, spacing=[0.7, 1.3])
    add_pd("3d_TV_L2_spacing", "ph3d", reg="TV", data="L2", alpha=0.05, L2=8, iterations=15, spacing=[0.7, 1.3, 2.1])
    add_pd("3d_TV_L2_L12", "ph32", reg="TV", data="L2", alpha=0.05, L2=12, iterations=100)
    add_pd("3d_TV_L2_xscale1", "ph3d", reg="TV", data="L2", alpha=0.05, L2=8, iterations=10, x_scale=1.0)
    # BASELINE config 1: full 256^2 Lena, TV-L2, alpha=0.05, 100 iterations, L2=8
    add_pd("c1_lena_TV_L2", "lena", reg="TV", data="L2", alpha=0.05, L2=8, iterations=100)
    # BASELINE config 2 (crop): Huber-L1 on salt & pepper, alpha=0.6, 200 iterations
    add_pd("c2_man_HUBER_L1", "man_sp", reg="HUBER", data="L1", alpha=0.6, L2=8, iterations=200)
    # iterates as the Observer sees them
    x, xs = ref_pd(bw_crop, reg="TV", data="L2", alpha=0.05, L2=8, iterations=6, iterates=True)
    pd_out["2d_TV_L2_iterates"] = xs
    pd_manifest["2d_TV_L2_iterates"] = dict(input="bw2d", reg="TV", data="L2", alpha=0.05, L2=8, iterations=6)
    np.savez_compressed(os.path.join(OUT, "pd.npz"), **pd_out)
    manifest["pd"] = pd_manifest

    # ------------------------------------------------------------ Tikhonov / ADMM
    ls_out = {}
    ls_manifest = {}

    def blurred(clean, var, spacing=None, poisson=False):
        dim = clean.ndim
        cov = var if dim == 1 else np.diag(var)
        A, _, _, _ = ref_deconv_ops(clean.shape, cov, spacing)
        n = noise.Noise(A(clean).reshape(clean.shape), seed=1)
        if poisson:
            n.add_poisson_noise(noise_level=0.05)
        else:
            n.add_gaussian_noise(noise_level=0.05)
        return n.get_noisy_data()

    ls_inputs = {
        "spike1d": blurred(spike, 1.5, poisson=True),                      # tests/solvers_test.py setup
        "bw2d": blurred(brainweb[40:104, 30:90].astype(np.float64), [1.5, 1.5], poisson=True),
        "lena64": blurred(lena512[200:264, 220:284].astype(np.float64), [1.0, 1.0]),
        "ph3d": blurred(phantom[16:40, 20:42, 18:38].astype(np.float64), [1.0, 1.0, 1.0]),
        "lena512": blurred(lena512.astype(np.float64), [1.0, 1.0]),
    }
    ls_var = {"spike1d": 1.5, "bw2d": [1.5, 1.5], "lena64": [1.0, 1.0], "ph3d": [1.0, 1.0, 1.0],
              "lena512": [1.0, 1.0]}
    for k, v in ls_inputs.items():
        if k != "lena512":
            ls_out["in/" + k] = v
    ls_out["in/lena512_f32"] = ls_inputs["lena512"].astype(np.float32)

    def add_admm(name, inp, alpha, rho, iterations, iter_max, x_scale=None, spacing=None):
        obs = ls_inputs[inp]
        dim = obs.ndim
        cov = ls_var[inp] if dim == 1 else np.diag(ls_var[inp])
        A, A_adj, D, D_adj = ref_deconv_ops(obs.shape, cov, spacing)
        b = obs.flatten()
        xs = float(np.max(obs)) if x_scale is None else x_scale
        s = admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=b, B=D, B_adj=D_adj, x0=obs.flatten(),
                                  dimension=dim, alpha=alpha, rho=rho, iterations=iterations,
                                  iter_max=iter_max, x_scale=xs)
        s.run()
        ls_out[name] = s.get_x()
        ls_manifest[name] = dict(kind="admm", input=inp, var=ls_var[inp], alpha=alpha, rho=rho,
                                 iterations=iterations, iter_max=iter_max, x_scale=x_scale, spacing=spacing)

    def add_tk(name, inp, alpha, iter_max, reg, x_scale=None):
        obs = ls_inputs[inp]
        dim = obs.ndim
        cov = ls_var[inp] if dim == 1 else np.diag(ls_var[inp])
        A, A_adj, D, D_adj = ref_deconv_ops(obs.shape, cov)
        ident = lambda x: x.flatten()
        b = obs.flatten()
        xs = float(np.max(obs)) if x_scale is None else x_scale
        s = tk.TikhonovLinearSolver(A=A, A_adj=A_adj, b=b, B=D if reg == "TK1" else ident,
                                    B_adj=D_adj if reg == "TK1" else ident, x0=obs.flatten(),
                                    alpha=alpha, iter_max=iter_max, x_scale=xs)
        s.run()
        ls_out[name] = s.get_x()
        ls_manifest[name] = dict(kind="tikhonov", input=inp, var=ls_var[inp], alpha=alpha,
                                 iter_max=iter_max, reg=reg, x_scale=x_scale)

    def add_pdd(name, inp, reg, alpha, iterations, iter_max, x_scale=None, L2=8):
        """default deconvolution wiring: ...interface.py:255-280 (TV) / :303-325 (Huber); tests/solvers_test.py:138-150"""
        obs = ls_inputs[inp]
        dim = obs.ndim
        cov = ls_var[inp] if dim == 1 else np.diag(ls_var[inp])
        A, A_adj, D, D_adj = ref_deconv_ops(obs.shape, cov)
        b = obs.flatten()
        x0 = obs.flatten()
        xs = float(np.max(obs)) if x_scale is None else x_scale
        s = pd.PrimalDualSolver(
            prox_f=lambda x, tau: prox.prox_linear_least_squares(x=x, tau=tau, A=A, A_adj=A_adj, b=b, x0=x0,
                                                                 iter_max=iter_max, x_scale=xs),
            prox_g_conj=prox.prox_tv_conj if reg == "TV" else prox.prox_huber_conj, B=D, B_conj=D_adj, L2=L2, x0=x0,
            alpha=alpha, iterations=iterations, x_scale=xs)
        s.run()
        ls_out[name] = s.get_x()
        ls_manifest[name] = dict(kind="pd_deconv", input=inp, var=ls_var[inp], reg=reg, alpha=alpha, iterations=iterations,
                                 iter_max=iter_max, x_scale=x_scale, L2=L2)

    add_pdd("pdd_1d_TV", "spike1d", "TV", 0.01, 10, 10)
    add_pdd("pdd_1d_TV_xs1", "spike1d", "TV", 0.01, 10, 10, x_scale=1.0)
    add_pdd("pdd_2d_TV", "bw2d", "TV", 0.01, 10, 10)
    add_pdd("pdd_2d_HUBER", "lena64", "HUBER", 0.02, 12, 8)
    add_pdd("pdd_3d_TV", "ph3d", "TV", 0.01, 6, 10)
    add_admm("admm_1d", "spike1d", 0.01, 0.5, 10, 10)
    add_admm("admm_1d_xs1", "spike1d", 0.01, 0.5, 10, 10, x_scale=1.0)
    add_admm("admm_2d", "bw2d", 0.01, 0.5, 10, 10)
    add_admm("admm_2d_c3crop", "lena64", 0.01, 0.1, 50, 10)
    add_admm("admm_3d", "ph3d", 0.01, 0.1, 8, 10)
    add_admm("admm_c3_lena512", "lena512", 0.01, 0.1, 5, 10)
    for reg in ("TK0", "TK1"):
        add_tk("tk_1d_%s" % reg, "spike1d", 0.01, 10, reg)
        add_tk("tk_2d_%s" % reg, "bw2d", 0.01, 10, reg)
        add_tk("tk_3d_%s" % reg, "ph3d", 0.05, 25, reg)
    # full-size result kept as float32 to bound the fixture size (used at 1e-6 only)
    ls_out["admm_c3_lena512"] = ls_out["admm_c3_lena512"].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "lsmr.npz"), **ls_out)
    manifest["lsmr"] = ls_manifest

    with open(os.path.join(OUT, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    for f in sorted(os.listdir(OUT)):
        print("%-16s %8.1f KiB" % (f, os.path.getsize(os.path.join(OUT, f)) / 1024.0))


if __name__ == "__main__":
    main()
