"""GPU parity tests (run with ``-m gpu`` on a B200).

Every test drives the CUDA path through the reference-shaped Python API, i.e.
through the C ABI of libnsol_b200.so, and compares with
  * the golden fixtures produced by the unmodified reference (tests/golden), and
  * the CPU oracle (oracle/nsol_oracle.py) on seeded inputs.

Bars (north_star): float64 primal-dual and stencils BIT-EXACT (stricter than the
1e-10 the north star asks for); float64 LSMR/ADMM <= 1e-10 relative max-abs;
float32 <= 1e-4 relative max-abs and PSNR/SSIM/NCC equal to three decimals.
"""
import numpy as np
import pytest

from conftest import rel_max
from oracle import nsol_oracle as orc

import nsol_b200.linear_operators as lo
import nsol_b200.primal_dual_solver as pd
import nsol_b200.admm_linear_solver as admm
import nsol_b200.tikhonov_linear_solver as tk
from nsol_b200 import _lib
from nsol_b200.proximal_operators import ProximalOperators as prox

pytestmark = pytest.mark.gpu

F64_LSMR_TOL = 1e-10     # north_star: fp64 within 1e-10 relative max-abs
F32_TOL = 1e-4           # north_star: fp32 within 1e-4 relative


def linops(dim, spacing=None):
    cls = getattr(lo, "LinearOperators%dD" % dim)
    if spacing is None:
        return cls()
    return cls(spacing=np.asarray(spacing, dtype=float) if dim > 1 else float(spacing[0]))


def make_pd(obs, reg, data, alpha, L2, iterations, alg_type="ALG2", spacing=None, x_scale=None, dtype=None):
    """Wired exactly like nsol/application/run_denoising.py:95-154."""
    dim = obs.ndim
    b = obs.flatten()
    x0 = obs.flatten()
    x_scale = float(np.max(obs)) if x_scale is None else x_scale
    grad, grad_adj = linops(dim, spacing).get_gradient_operators()
    X_shape = obs.shape
    Z_shape = (dim * obs.shape[0],) + obs.shape[1:]
    D_1D = lambda x: grad(x.reshape(*X_shape)).flatten()
    D_adj_1D = lambda x: grad_adj(x.reshape(*Z_shape)).flatten()
    if data == "L1":
        prox_f = lambda x, tau: prox.prox_ell1_denoising(x, tau, x0=b, x_scale=x_scale)
    else:
        prox_f = lambda x, tau: prox.prox_ell2_denoising(x, tau, x0=b, x_scale=x_scale)
    prox_g_conj = {"TV": prox.prox_tv_conj, "HUBER": prox.prox_huber_conj, "TK1": lambda q, s: q / (1 + s)}[reg]
    return pd.PrimalDualSolver(prox_f=prox_f, prox_g_conj=prox_g_conj, B=D_1D, B_conj=D_adj_1D, L2=L2, x0=x0,
                               alpha=alpha, iterations=iterations, x_scale=x_scale, alg_type=alg_type, dtype=dtype)


def run_pd(obs, **kw):
    s = make_pd(obs, **kw)
    s.run()
    return s.get_x()


def assert_measures_3dec(x, ref, clean):
    for f in (orc.psnr, orc.ncc, orc.ssim_1d):
        a, b = f(x, clean), f(ref, clean)
        assert round(a, 3) == round(b, 3), (f.__name__, a, b)


# ------------------------------------------------------------------ operators
def test_gradient_operators_bit_exact(golden):
    for name, meta in sorted(golden.manifest["ops"].items()):
        if not name.startswith("g"):
            continue
        x = golden("ops", name + "/x")
        y = golden("ops", name + "/y")
        ops = linops(x.ndim, meta["spacing"])
        grad, grad_adj = ops.get_gradient_operators()
        g = grad(x)
        assert g.shape == golden("ops", name + "/grad").shape
        assert np.array_equal(g, golden("ops", name + "/grad")), name
        assert np.array_equal(grad_adj(y), golden("ops", name + "/grad_adj")), name
        # grad == concat(dx, dy, dz), grad_adj == sum d*_adj  (tests/kernels_test.py:222-301)
        parts, adj = [], None
        blocks = np.array_split(y, x.ndim)
        for k, nm in enumerate(["dx", "dy", "dz"][:x.ndim]):
            D, D_adj = getattr(ops, "get_%s_operators" % nm)()
            parts.append(D(x))
            adj = D_adj(blocks[k]) if adj is None else adj + D_adj(blocks[k])
        assert np.array_equal(np.concatenate(parts), g), name
        assert np.array_equal(adj, golden("ops", name + "/grad_adj")), name


def test_blur_operators(golden):
    for name, meta in sorted(golden.manifest["ops"].items()):
        if not name.startswith("b"):
            continue
        x = golden("ops", name + "/x")
        dim = x.ndim
        cov = meta["var"] if dim == 1 else np.diag(meta["var"])
        ops = linops(dim, meta["spacing"])
        A, A_adj = ops.get_gaussian_blurring_operators(cov)
        ref = golden("ops", name + "/A")
        assert rel_max(A(x), ref) < 1e-14, name
        assert rel_max(A_adj(x), ref) < 1e-14, name
        # dense path with the same mask
        Ad, _ = ops.get_convolution_and_adjoint_convolution_operators(golden("ops", name + "/kernel"))
        Ad.taps = None
        assert rel_max(Ad(x), ref) < 1e-14, name


def test_adjointness_properties():
    """tests/kernels_test.py:138-335 on the CUDA operators: <Ax,y> == <x,A'y> to 1e-10."""
    rng = np.random.RandomState(0)
    for shape in [(50,), (50, 50), (50, 50, 10)]:
        dim = len(shape)
        spacing = rng.rand(dim) + 0.5
        ops = linops(dim, spacing)
        x = rng.rand(*shape)
        grad, grad_adj = ops.get_gradient_operators()
        y = rng.rand(*((dim * shape[0],) + shape[1:]))
        assert round(abs(np.sum(grad(x) * y) - np.sum(x * grad_adj(y))), 10) == 0
        A, A_adj = ops.get_gaussian_blurring_operators(1.5 if dim == 1 else np.eye(dim) * 1.5)
        z = rng.rand(*shape)
        assert round(abs(np.sum(A(x) * z) - np.sum(x * A_adj(z))), 10) == 0


def test_standalone_prox_maps():
    rng = np.random.RandomState(5)
    x = rng.randn(1000) * 2
    x0 = rng.rand(1000) * 3
    assert np.array_equal(prox.prox_tv_conj(x, 0.3), orc.prox_tv_conj(x, 0.3))
    assert np.array_equal(prox.prox_huber_conj(np.array(x), 0.3), orc.prox_huber_conj(x, 0.3))
    assert np.array_equal(prox.prox_ell1_denoising(x, 0.4, x0, 1.7), orc.prox_ell1_denoising(x, 0.4, x0, 1.7))
    assert np.array_equal(prox.prox_ell2_denoising(x, 0.4, x0, 1.7), orc.prox_ell2_denoising(x, 0.4, x0, 1.7))


# ------------------------------------------------------------------ primal-dual vs reference fixtures
def test_primal_dual_fp64_bit_exact_vs_reference(golden):
    for name, meta in sorted(golden.manifest["pd"].items()):
        if name.endswith("_iterates"):
            continue
        obs = golden("pd", "in/" + meta["input"])
        x = run_pd(obs, reg=meta["reg"], data=meta["data"], alpha=meta["alpha"], L2=meta["L2"],
                   iterations=meta["iterations"], alg_type=meta.get("alg_type", "ALG2"),
                   spacing=meta.get("spacing"), x_scale=meta.get("x_scale"))
        ref = golden("pd", name)
        assert np.array_equal(x, ref), (name, rel_max(x, ref))


def test_primal_dual_fp32_vs_reference(golden):
    for name in ("c1_lena_TV_L2", "c2_man_HUBER_L1", "3d_TV_L2_L12", "3d_HUBER_L1", "2d_TK1_L2", "1d_TV_L2"):
        meta = golden.manifest["pd"][name]
        obs = golden("pd", "in/" + meta["input"])
        x = run_pd(obs, reg=meta["reg"], data=meta["data"], alpha=meta["alpha"], L2=meta["L2"],
                   iterations=meta["iterations"], dtype="float32")
        ref = golden("pd", name)
        assert rel_max(x, ref) < F32_TOL, (name, rel_max(x, ref))
        assert_measures_3dec(x, ref, obs.reshape(-1))


def test_primal_dual_observer_iterates(golden):
    meta = golden.manifest["pd"]["2d_TV_L2_iterates"]
    obs = golden("pd", "in/" + meta["input"])
    from nsol_b200.observer import Observer
    s = make_pd(obs, reg="TV", data="L2", alpha=meta["alpha"], L2=meta["L2"], iterations=meta["iterations"])
    o = Observer()
    s.set_observer(o)
    s.run()
    assert np.array_equal(np.array(o.get_x_list()), golden("pd", "2d_TV_L2_iterates"))
    assert s.get_computational_time().total_seconds() > 0


# ------------------------------------------------------------------ primal-dual vs oracle on seeded inputs
@pytest.mark.parametrize("shape", [(1,), (2,), (333,), (1, 1), (1, 7), (9, 1), (3, 2), (67, 258), (130, 131),
                                   (1, 1, 1), (2, 3, 1), (1, 4, 6), (5, 1, 6), (19, 21, 130), (33, 18, 66), (9, 40, 260)])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_primal_dual_ragged_shapes_vs_oracle(shape, dtype):
    rng = np.random.RandomState(sum(shape))
    obs = rng.rand(*shape) * 255
    ctx = _lib.context()
    ctx.set_tuning("pd_zc", 8)     # several z-chunks even for small volumes
    ctx.set_tuning("pd_ty", 4)
    ctx.set_tuning("pd_variant", 2 if sum(shape) % 2 else 1)   # exercise both kernel variants
    try:
        for reg, data, alpha in (("TV", "L2", 0.05), ("HUBER", "L1", 0.6)):
            x = run_pd(obs, reg=reg, data=data, alpha=alpha, L2=8, iterations=12, dtype=dtype)
            ref = orc.primal_dual_denoise(obs.reshape(-1), shape, reg=reg, data=data, alpha=alpha, L2=8, iterations=12,
                                          x_scale=float(obs.max()))
            if dtype == "float64":
                assert np.array_equal(x, ref), (shape, reg, rel_max(x, ref))
            else:
                assert rel_max(x, ref) < F32_TOL, (shape, reg, rel_max(x, ref))
    finally:
        ctx.set_tuning("pd_zc", 0)
        ctx.set_tuning("pd_ty", 0)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_primal_dual_tiling_and_variant_independent(dtype):
    """idempotence property: the result does not depend on z-chunk length, rows per CTA or on
    the kernel variant (1 = register-pipelined LDG loads, 2 = TMA bulk-async staged tiles)."""
    rng = np.random.RandomState(11)
    obs = rng.rand(70, 45, 136) * 255
    ctx = _lib.context()
    outs = []
    try:
        for variant in (1, 2):
            for zc, ty in ((0, 0), (8, 4), (16, 8), (64, 2), (3, 1)):
                ctx.set_tuning("pd_variant", variant)
                ctx.set_tuning("pd_zc", zc)
                ctx.set_tuning("pd_ty", ty)
                for reg, data in (("TV", "L2"), ("HUBER", "L1")):
                    outs.append((reg, run_pd(obs, reg=reg, data=data, alpha=0.05 if data == "L2" else 0.6, L2=12,
                                             iterations=10, dtype=dtype)))
    finally:
        for key in ("pd_variant", "pd_zc", "pd_ty"):
            ctx.set_tuning(key, 0)
    for reg, o in outs[2:]:
        ref = outs[0][1] if reg == "TV" else outs[1][1]
        if dtype == "float64":
            assert np.array_equal(o, ref)
        else:
            assert rel_max(o, ref) < 1e-5


def test_primal_dual_sweep_batched_vs_oracle(golden):
    """BASELINE config 5 shape of work: alpha sweep batched into one launch per iteration."""
    obs = golden("pd", "in/man_sp")
    alphas = np.linspace(0.001, 0.05, 7)
    for reg in ("TV", "HUBER", "TK1"):
        s = make_pd(obs, reg=reg, data="L2", alpha=0.01, L2=8, iterations=40)
        xs = s.run_sweep(alphas)
        for i, a in enumerate(alphas):
            ref = orc.primal_dual_denoise(obs.reshape(-1), obs.shape, reg=reg, data="L2", alpha=a, L2=8, iterations=40,
                                          x_scale=float(obs.max()))
            assert np.array_equal(xs[i], ref), (reg, a)


def test_x_scale_invariance_pd():
    """tests/solvers_test.py:102-352 property for the denoising wiring."""
    rng = np.random.RandomState(2)
    obs = rng.rand(40, 37) * 200 + 1
    xs = float(obs.max())
    r1 = run_pd(obs / xs, reg="TV", data="L2", alpha=0.05, L2=8, iterations=20, x_scale=1.0)
    r2 = run_pd(obs, reg="TV", data="L2", alpha=0.05, L2=8, iterations=20, x_scale=xs)
    assert round(np.linalg.norm(r2 - r1 * xs), 7) == 0


def test_full_size_512_cube_locality_vs_oracle():
    """BASELINE config 4 size (512^3): after k iterations a voxel depends on a radius-k
    neighbourhood, so corners/faces of the full volume can be checked against the oracle run
    on a crop with a (k+1)-voxel margin -- including the zero boundary on the low/high sides."""
    n, k, core, m = 512, 3, 20, 5
    rng = np.random.RandomState(1)
    obs = rng.rand(n, n, n).astype(np.float64) * 255
    xs = float(obs.max())
    x = run_pd(obs, reg="TV", data="L2", alpha=0.05, L2=8, iterations=k, x_scale=xs).reshape(n, n, n)
    x32 = run_pd(obs, reg="TV", data="L2", alpha=0.05, L2=8, iterations=k, x_scale=xs, dtype="float32").reshape(n, n, n)
    for lo_side in (True, False):
        sl = slice(0, core + m) if lo_side else slice(n - core - m, n)
        inner = slice(0, core) if lo_side else slice(m, core + m)
        crop = np.ascontiguousarray(obs[sl, sl, sl])
        ref = orc.primal_dual_denoise(crop.reshape(-1), crop.shape, reg="TV", data="L2", alpha=0.05, L2=8,
                                      iterations=k, x_scale=xs).reshape(crop.shape)
        got = x[sl, sl, sl]
        assert np.array_equal(got[inner, inner, inner], ref[inner, inner, inner])
        assert rel_max(x32[sl, sl, sl][inner, inner, inner], ref[inner, inner, inner]) < F32_TOL
    # interior block straddling tile (x: 64/128), row (y: 8) and chunk (z) seams
    sl = slice(240, 240 + core + 2 * m)
    crop = np.ascontiguousarray(obs[sl, sl, sl])
    ref = orc.primal_dual_denoise(crop.reshape(-1), crop.shape, reg="TV", data="L2", alpha=0.05, L2=8,
                                  iterations=k, x_scale=xs).reshape(crop.shape)
    inner = slice(m, m + core)
    assert np.array_equal(x[sl, sl, sl][inner, inner, inner], ref[inner, inner, inner])


# ------------------------------------------------------------------ LSMR / Tikhonov / ADMM
def deconv_callables(shape, var, spacing=None):
    """nsol/application/run_deconvolution.py:109-129."""
    dim = len(shape)
    ops = linops(dim, spacing)
    cov = var if dim == 1 else np.diag(var)
    A, A_adj = ops.get_gaussian_blurring_operators(cov)
    grad, grad_adj = ops.get_gradient_operators()
    Z_shape = (dim * shape[0],) + tuple(shape[1:])
    return (lambda x: A(x.reshape(*shape)).flatten(), lambda x: A_adj(x.reshape(*shape)).flatten(),
            lambda x: grad(x.reshape(*shape)).flatten(), lambda x: grad_adj(x.reshape(*Z_shape)).flatten())


@pytest.mark.parametrize("name", ["admm_1d", "admm_1d_xs1", "admm_2d", "admm_2d_c3crop", "admm_3d"])
def test_admm_vs_reference(golden, name):
    meta = golden.manifest["lsmr"][name]
    obs = golden("lsmr", "in/" + meta["input"])
    A, A_adj, D, D_adj = deconv_callables(obs.shape, meta["var"], meta["spacing"])
    xs = meta["x_scale"] or float(obs.max())
    s = admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(), dimension=obs.ndim,
                              alpha=meta["alpha"], rho=meta["rho"], iterations=meta["iterations"],
                              iter_max=meta["iter_max"], x_scale=xs)
    s.run()
    ref = golden("lsmr", name)
    assert rel_max(s.get_x(), ref) < F64_LSMR_TOL, (name, rel_max(s.get_x(), ref))
    s32 = admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(), dimension=obs.ndim,
                                alpha=meta["alpha"], rho=meta["rho"], iterations=meta["iterations"],
                                iter_max=meta["iter_max"], x_scale=xs, dtype="float32")
    s32.run()
    assert rel_max(s32.get_x(), ref) < F32_TOL, (name, rel_max(s32.get_x(), ref))
    assert_measures_3dec(s32.get_x(), ref, obs.reshape(-1))      # PSNR / NCC / SSIM against the observation


def test_admm_config3_lena512(golden):
    """BASELINE config 3 at full size (5 outer x 10 LSMR iterations; float32 fixture -> 1e-6)."""
    meta = golden.manifest["lsmr"]["admm_c3_lena512"]
    obs = golden("lsmr", "in/lena512_f32").astype(np.float64)
    A, A_adj, D, D_adj = deconv_callables(obs.shape, meta["var"])
    s = admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(), dimension=2,
                              alpha=meta["alpha"], rho=meta["rho"], iterations=meta["iterations"],
                              iter_max=meta["iter_max"], x_scale=float(obs.max()))
    s.run()
    # the stored input was rounded to float32, so compare with the oracle on the same rounded input
    Ao, Ao_adj, Do, Do_adj = orc.deconvolution_operators(obs.shape, np.diag(meta["var"]))
    ref = orc.admm_tv(Ao, Ao_adj, Do, Do_adj, obs.reshape(-1), obs.reshape(-1), 2, alpha=meta["alpha"], rho=meta["rho"],
                      iterations=meta["iterations"], iter_max=meta["iter_max"], x_scale=float(obs.max()))
    assert rel_max(s.get_x(), ref) < F64_LSMR_TOL
    assert rel_max(s.get_x(), golden("lsmr", "admm_c3_lena512")) < 1e-4


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("shape,nslabs", [((40, 36), 1), ((40, 36), 2), ((41, 36), 3), ((18, 10, 20), 2), ((22, 6, 20), 3)])
def test_admm_zslab_emulation_matches_unsharded(shape, nslabs, dtype):
    """the ADMM / LSMR path z-slab sharded.  S slabs are driven phase by phase in one process
    (nsol_lsmr_slab_phase); the halo exchange (ring for the periodic blur, neighbours for the gradient) is
    done with device copies and the scalar all-reduce on the host -- the same step program
    (distributed.slab_admm_program) the NCCL driver SlabADMM executes on S GPUs.  Must agree with the
    unsharded solver (only the order of the norm reductions differs)."""
    from nsol_b200.distributed import SlabLsq, slab_admm_program, run_slab_admm_emulated, slab_bounds
    from nsol_b200.linear_solver import probe_least_squares
    rng = np.random.RandomState(17)
    obs = rng.rand(*shape) * 200 + 20
    var = [1.0] * len(shape)
    alpha, rho, iterations, iter_max = 0.02, 0.2, 4, 6
    xs = float(obs.max())
    A, A_adj, D, D_adj = deconv_callables(shape, var)
    ref_solver = admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(), dimension=len(shape),
                                       alpha=alpha, rho=rho, iterations=iterations, iter_max=iter_max, x_scale=xs, dtype=dtype)
    ref_solver.run()
    ref = ref_solver.get_x()
    info = probe_least_squares(A, A_adj, D, D_adj, obs.size)
    ctx = _lib.context()
    slabs = []
    try:
        for r in range(nslabs):
            z_lo, z_hi = slab_bounds(shape[0], r, nslabs)
            sl = SlabLsq(ctx, dict(info, shape=(z_hi - z_lo,) + tuple(shape[1:])), dtype, r, nslabs)
            part = np.ascontiguousarray(obs[z_lo:z_hi]).reshape(-1) / xs
            sl.upload(part, part)
            slabs.append(sl)
        run_slab_admm_emulated(slabs, slab_admm_program(iterations, iter_max, alpha, rho))
        out = np.concatenate([sl.download(xs) for sl in slabs])
    finally:
        for sl in slabs:
            sl.close()
    tol = 1e-9 if dtype == "float64" else F32_TOL
    assert rel_max(out, ref) < tol, rel_max(out, ref)


@pytest.mark.parametrize("name", ["tk_1d_TK0", "tk_1d_TK1", "tk_2d_TK0", "tk_2d_TK1", "tk_3d_TK0", "tk_3d_TK1"])
def test_tikhonov_vs_reference(golden, name):
    meta = golden.manifest["lsmr"][name]
    obs = golden("lsmr", "in/" + meta["input"])
    A, A_adj, D, D_adj = deconv_callables(obs.shape, meta["var"])
    ident = lambda x: x.flatten()
    B, B_adj = (D, D_adj) if meta["reg"] == "TK1" else (ident, ident)
    s = tk.TikhonovLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=B, B_adj=B_adj, x0=obs.flatten(), alpha=meta["alpha"],
                                iter_max=meta["iter_max"], x_scale=float(obs.max()))
    s.run()
    assert rel_max(s.get_x(), golden("lsmr", name)) < F64_LSMR_TOL, (name, rel_max(s.get_x(), golden("lsmr", name)))


def test_x_scale_invariance_admm_and_tikhonov(golden):
    """tests/solvers_test.py:102-224 (1-D case), default hyper-parameters, accuracy 1e-7."""
    obs = golden("lsmr", "in/spike1d")
    xs = float(obs.max())
    A, A_adj, D, D_adj = deconv_callables(obs.shape, 1.5)
    res = {}
    for tag, data, scale in (("scaled", obs / xs, 1.0), ("raw", obs, xs)):
        t = tk.TikhonovLinearSolver(A=A, A_adj=A_adj, B=D, B_adj=D_adj, b=data.flatten(), x0=data.flatten(), x_scale=scale)
        t.run()
        a = admm.ADMMLinearSolver(A=A, A_adj=A_adj, B=D, B_adj=D_adj, b=data.flatten(), x0=data.flatten(), x_scale=scale,
                                  dimension=1)
        a.run()
        res[tag] = (t.get_x(), a.get_x())
    for i in range(2):
        assert round(np.linalg.norm(res["raw"][i] - res["scaled"][i] * xs), 7) == 0


def test_lsmr_edge_cases():
    """zero right-hand side (lsmr.py:307-315) and maxiter = 0 return x = 0."""
    shape = (12, 10)
    A, A_adj, D, D_adj = deconv_callables(shape, [1.0, 1.0])
    z = np.zeros(120)
    t = tk.TikhonovLinearSolver(A=A, A_adj=A_adj, B=D, B_adj=D_adj, b=z, x0=z)
    t.run()
    assert np.array_equal(t.get_x(), z)
    b = np.random.RandomState(0).rand(120)
    t = tk.TikhonovLinearSolver(A=A, A_adj=A_adj, B=D, B_adj=D_adj, b=b, x0=b, iter_max=0)
    t.run()
    assert np.array_equal(t.get_x(), z)


# ------------------------------------------------------------------ z-slab decomposition (single-GPU emulation)
@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("nslabs", [2, 3])
def test_zslab_emulation_matches_unsharded(dtype, nslabs, split):
    """run S z-slabs one after another on one GPU with explicit halo planes
    (exactly the buffers the NCCL exchange fills on S GPUs) and compare with the unsharded run
    bit for bit."""
    import ctypes as C
    from nsol_b200.distributed import slab_bounds
    rng = np.random.RandomState(3)
    shape = (23, 10, 68)
    iters = 9
    obs = rng.rand(*shape) * 255
    xs = float(obs.max())
    ref = run_pd(obs, reg="TV", data="L2", alpha=0.05, L2=8, iterations=iters, x_scale=xs, dtype=dtype)

    ctx = _lib.context()
    lib = ctx.lib
    dcode = _lib.dtype_code(dtype)
    esz = 4 if dcode == _lib.F32 else 8
    plane = shape[1] * shape[2]
    alpha = np.array([0.05])
    plans, halos, spans = [], [], []
    if split:
        ctx.set_tuning("pd_zc", 2)      # >= 3 z-chunks per slab: the overlapped (split) iteration
    for r in range(nslabs):
        z_lo, z_hi = slab_bounds(shape[0], r, nslabs)
        spans.append((z_lo, z_hi))
        desc = _lib.PdDesc()
        desc.grid = _lib.make_grid((z_hi - z_lo,) + shape[1:], None, dcode, 1)
        desc.reg, desc.data, desc.alg = _lib.REG["TV"], _lib.DATA["L2"], _lib.ALG["ALG2"]
        desc.huber_gamma, desc.L2 = 0.05, 8.0
        desc.x_scale = desc.x0_scale = desc.b_scale = xs
        desc.alpha = alpha.ctypes.data_as(_lib.c_double_p)
        h = C.c_void_p()
        ctx.check(lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(h)))
        plans.append(h)
        bufs = {k: ctx.device_alloc(plane * esz) for k in ("above", "below", "pz_below")}
        halos.append(bufs)
        ctx.check(lib.nsol_pd_plan_set_halo(h, bufs["above"].ptr if r < nslabs - 1 else None,
                                            bufs["below"].ptr if r > 0 else None,
                                            bufs["pz_below"].ptr if r > 0 else None))
        slab = np.ascontiguousarray(obs[z_lo:z_hi]).reshape(-1)
        ctx.check(lib.nsol_pd_plan_reset_host(h, slab.ctypes.data, None, None))
    def exchange(upcoming):
        fn = lib.nsol_pd_plan_boundary_planes_next if upcoming else lib.nsol_pd_plan_boundary_planes
        bounds = []
        for h in plans:
            a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
            ctx.check(fn(h, C.byref(a), C.byref(b), C.byref(c)))
            bounds.append((a, b, c))
        for r in range(nslabs):
            if r < nslabs - 1:
                ctx.check(lib.nsol_memcpy_d2d(ctx.handle, halos[r]["above"].ptr, bounds[r + 1][0], plane * esz, None))
            if r > 0:
                ctx.check(lib.nsol_memcpy_d2d(ctx.handle, halos[r]["below"].ptr, bounds[r - 1][1], plane * esz, None))
                ctx.check(lib.nsol_memcpy_d2d(ctx.handle, halos[r]["pz_below"].ptr, bounds[r - 1][2], plane * esz, None))

    try:
        if split:
            # the flow of SlabPrimalDual.iterate(overlap=True): boundary chunks, exchange of the NEW
            # boundary planes (overlapped with the interior chunks on real GPUs), interior chunks
            assert all(lib.nsol_pd_plan_chunks(h) >= 3 for h in plans)
            exchange(False)
            for _ in range(iters):
                for h in plans:
                    ctx.check(lib.nsol_pd_plan_iterate_part(h, 1, None))
                exchange(True)
                for h in plans:
                    ctx.check(lib.nsol_pd_plan_iterate_part(h, 2, None))
        for _ in range(0 if split else iters):
            bounds = []
            for h in plans:
                a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
                ctx.check(lib.nsol_pd_plan_boundary_planes(h, C.byref(a), C.byref(b), C.byref(c)))
                bounds.append((a, b, c))
            for r in range(nslabs):
                if r < nslabs - 1:      # from the upper neighbour: its first xbar plane
                    ctx.check(lib.nsol_memcpy_d2d(ctx.handle, halos[r]["above"].ptr, bounds[r + 1][0], plane * esz, None))
                if r > 0:               # from the lower neighbour: its last xbar and p_z planes
                    ctx.check(lib.nsol_memcpy_d2d(ctx.handle, halos[r]["below"].ptr, bounds[r - 1][1], plane * esz, None))
                    ctx.check(lib.nsol_memcpy_d2d(ctx.handle, halos[r]["pz_below"].ptr, bounds[r - 1][2], plane * esz, None))
            for h in plans:
                ctx.check(lib.nsol_pd_plan_iterate(h, 1, None))
        parts = []
        for (z_lo, z_hi), h in zip(spans, plans):
            out = np.empty((z_hi - z_lo) * plane)
            ctx.check(lib.nsol_pd_plan_get_x_host(h, out.ctypes.data, None))
            parts.append(out)
    finally:
        ctx.set_tuning("pd_zc", 0)
        for h in plans:
            lib.nsol_pd_plan_destroy(h)
    if dtype == "float64":
        assert np.array_equal(np.concatenate(parts), ref)
    else:
        assert rel_max(np.concatenate(parts), ref) < 1e-5


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("nslabs,zc", [(2, 0), (3, 2)])
def test_zslab_link_emulation_matches_unsharded(dtype, nslabs, zc, variant):
    """The in-kernel halo exchange (nsol_pd_plan_link_*): S slabs on ONE GPU whose link blocks are
    connected by raw pointers (on S GPUs: CUDA IPC handles), driven alternately on one stream --
    the kernels push their boundary planes into the neighbour's receive slots and raise / wait on
    the generation flags themselves.  Two solves back to back exercise the generation counters
    across a reset.  Bit-exact against the unsharded run."""
    import ctypes as C
    from nsol_b200.distributed import slab_bounds
    rng = np.random.RandomState(5)
    shape = (23, 10, 68)
    obs = rng.rand(*shape) * 255
    xs = float(obs.max())
    ctx = _lib.context()
    lib = ctx.lib
    dcode = _lib.dtype_code(dtype)
    plane = shape[1] * shape[2]
    alpha = np.array([0.05])
    ctx.set_tuning("pd_variant", variant)
    ctx.set_tuning("pd_zc", zc)          # zc = 2: >= 3 chunks per slab, boundary chunks scheduled first
    plans, spans, blocks = [], [], []
    try:
        for r in range(nslabs):
            z_lo, z_hi = slab_bounds(shape[0], r, nslabs)
            spans.append((z_lo, z_hi))
            desc = _lib.PdDesc()
            desc.grid = _lib.make_grid((z_hi - z_lo,) + shape[1:], None, dcode, 1)
            desc.reg, desc.data, desc.alg = _lib.REG["TV"], _lib.DATA["L2"], _lib.ALG["ALG2"]
            desc.huber_gamma, desc.L2 = 0.05, 8.0
            desc.x_scale = desc.x0_scale = desc.b_scale = xs
            desc.alpha = alpha.ctypes.data_as(_lib.c_double_p)
            h = C.c_void_p()
            ctx.check(lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(h)))
            plans.append(h)
            blk, nbytes = C.c_void_p(), C.c_size_t()
            ctx.check(lib.nsol_pd_plan_link_create(h, C.byref(blk), C.byref(nbytes)))
            assert nbytes.value >= 6 * plane * (4 if dcode == _lib.F32 else 8)
            blocks.append(blk)
        for r, h in enumerate(plans):
            ctx.check(lib.nsol_pd_plan_link_connect(h, blocks[r - 1] if r > 0 else None,
                                                    blocks[r + 1] if r < nslabs - 1 else None))
        for iters in (7, 4):            # second solve: generations continue across the reset
            ref = run_pd(obs, reg="TV", data="L2", alpha=0.05, L2=8, iterations=iters, x_scale=xs, dtype=dtype)
            ctx.set_tuning("pd_variant", variant)
            ctx.set_tuning("pd_zc", zc)
            for (z_lo, z_hi), h in zip(spans, plans):
                slab = np.ascontiguousarray(obs[z_lo:z_hi]).reshape(-1)
                ctx.check(lib.nsol_pd_plan_reset_host(h, slab.ctypes.data, None, None))
            for h in plans:
                ctx.check(lib.nsol_pd_plan_iterate(h, 0, None))      # publish the start state
            for _ in range(iters):
                for h in plans:
                    ctx.check(lib.nsol_pd_plan_iterate(h, 1, None))
            parts = []
            for (z_lo, z_hi), h in zip(spans, plans):
                out = np.empty((z_hi - z_lo) * plane)
                ctx.check(lib.nsol_pd_plan_get_x_host(h, out.ctypes.data, None))   # also checks the link status
                parts.append(out)
            if dtype == "float64":
                assert np.array_equal(np.concatenate(parts), ref)
            else:
                assert rel_max(np.concatenate(parts), ref) < 1e-5
    finally:
        ctx.set_tuning("pd_zc", 0)
        ctx.set_tuning("pd_variant", 0)
        for h in plans:
            lib.nsol_pd_plan_destroy(h)


def test_zslab_link_timeout_is_reported():
    """A linked plan whose neighbour never publishes must not hang: the wait gives up after
    link_timeout_ms (50 ms here, 5 s by default) and get_x / link_status report it."""
    import ctypes as C
    ctx = _lib.context()
    lib = ctx.lib
    ctx.set_tuning("link_timeout_ms", 50)
    shape = (6, 4, 8)
    obs = np.random.RandomState(0).rand(*shape)
    alpha = np.array([0.05])
    plans, blocks = [], []
    try:
        for r in range(2):
            desc = _lib.PdDesc()
            desc.grid = _lib.make_grid(shape, None, _lib.F64, 1)
            desc.reg, desc.data, desc.alg = _lib.REG["TV"], _lib.DATA["L2"], _lib.ALG["ALG2"]
            desc.huber_gamma, desc.L2 = 0.05, 8.0
            desc.x_scale = desc.x0_scale = desc.b_scale = 1.0
            desc.alpha = alpha.ctypes.data_as(_lib.c_double_p)
            h = C.c_void_p()
            ctx.check(lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(h)))
            plans.append(h)
            blk = C.c_void_p()
            ctx.check(lib.nsol_pd_plan_link_create(h, C.byref(blk), None))
            blocks.append(blk)
        ctx.check(lib.nsol_pd_plan_link_connect(plans[0], None, blocks[1]))
        ctx.check(lib.nsol_pd_plan_link_connect(plans[1], blocks[0], None))
        flat = np.ascontiguousarray(obs.reshape(-1))
        ctx.check(lib.nsol_pd_plan_reset_host(plans[0], flat.ctypes.data, None, None))
        ctx.check(lib.nsol_pd_plan_iterate(plans[0], 1, None))       # plan 1 never publishes
        with pytest.raises(RuntimeError, match="timed out"):
            ctx.check(lib.nsol_pd_plan_link_status(plans[0], None))
    finally:
        ctx.set_tuning("link_timeout_ms", 0)
        for h in plans:
            lib.nsol_pd_plan_destroy(h)


# ------------------------------------------------------------------ primal-dual deconvolution (prox_linear_least_squares)
def make_pd_deconv(obs, var, reg, alpha, iterations, iter_max, x_scale, L2=8, dtype=None):
    """nsol/deconvolution_solver_parameter_study_interface.py:255-280 (TV) / :303-325 (Huber)."""
    A, A_adj, D, D_adj = deconv_callables(obs.shape, var)
    b = obs.flatten()
    x0 = obs.flatten()
    return pd.PrimalDualSolver(
        prox_f=lambda x, tau: prox.prox_linear_least_squares(x=x, tau=tau, A=A, A_adj=A_adj, b=b, x0=x0,
                                                             iter_max=iter_max, x_scale=x_scale),
        prox_g_conj=prox.prox_tv_conj if reg == "TV" else prox.prox_huber_conj, B=D, B_conj=D_adj, L2=L2, x0=x0,
        alpha=alpha, iterations=iterations, x_scale=x_scale, dtype=dtype)


@pytest.mark.parametrize("name", ["pdd_1d_TV", "pdd_1d_TV_xs1", "pdd_2d_TV", "pdd_2d_HUBER", "pdd_3d_TV"])
def test_primal_dual_deconvolution_vs_reference(golden, name):
    meta = golden.manifest["lsmr"][name]
    obs = golden("lsmr", "in/" + meta["input"])
    xs = meta["x_scale"] or float(obs.max())
    s = make_pd_deconv(obs, meta["var"], meta["reg"], meta["alpha"], meta["iterations"], meta["iter_max"], xs, meta["L2"])
    s.run()
    ref = golden("lsmr", name)
    assert rel_max(s.get_x(), ref) < F64_LSMR_TOL, (name, rel_max(s.get_x(), ref))
    s32 = make_pd_deconv(obs, meta["var"], meta["reg"], meta["alpha"], meta["iterations"], meta["iter_max"], xs, meta["L2"],
                         dtype="float32")
    s32.run()
    assert rel_max(s32.get_x(), ref) < F32_TOL, (name, rel_max(s32.get_x(), ref))
    assert_measures_3dec(s32.get_x(), ref, obs.reshape(-1))


def test_x_scale_invariance_pd_deconvolution(golden):
    """tests/solvers_test.py:138-150, 193-224 (Primal-Dual part), accuracy 1e-7."""
    obs = golden("lsmr", "in/spike1d")
    xs = float(obs.max())
    r1 = make_pd_deconv(obs / xs, 1.5, "TV", 0.01, 10, 10, 1.0)
    r1.run()
    r2 = make_pd_deconv(obs, 1.5, "TV", 0.01, 10, 10, xs)
    r2.run()
    assert round(np.linalg.norm(r2.get_x() - r1.get_x() * xs), 7) == 0


def test_standalone_prox_linear_least_squares(golden):
    """prox_linear_least_squares called on arrays is a Tikhonov/LSMR solve (proximal_operators.py:44-78)."""
    obs = golden("lsmr", "in/bw2d")
    A, A_adj, _, _ = deconv_callables(obs.shape, [1.5, 1.5])
    Ao, Ao_adj, _, _ = orc.deconvolution_operators(obs.shape, np.diag([1.5, 1.5]))
    xs = float(obs.max())
    x = obs.flatten() / xs * 0.9
    got = prox.prox_linear_least_squares(x, 0.3, A, A_adj, obs.flatten(), obs.flatten(), iter_max=10, x_scale=xs)
    ident = lambda v: v.reshape(-1)
    ref = orc.tikhonov_lsmr(Ao, Ao_adj, ident, ident, obs.flatten() / xs, obs.flatten() / xs, alpha=1 / 0.3, b_reg=x,
                            iter_max=10, x_scale=xs)
    assert rel_max(got, ref) < F64_LSMR_TOL


@pytest.mark.parametrize("path", [1, 2, 3, 4])
def test_lsmr_multi_kernel_and_cooperative_paths(golden, path):
    """Both LSMR implementations (1 = one kernel per phase + CUDA graph for ADMM, 2 = a single
    cooperative launch with grid.sync between phases) against the reference fixtures."""
    ctx = _lib.context()
    ctx.set_tuning("lsmr_path", path)
    try:
        for name in ("admm_1d", "admm_2d", "admm_2d_c3crop", "admm_3d", "tk_2d_TK0", "tk_3d_TK1", "pdd_2d_TV"):
            meta = golden.manifest["lsmr"][name]
            obs = golden("lsmr", "in/" + meta["input"])
            xs = meta.get("x_scale") or float(obs.max())
            A, A_adj, D, D_adj = deconv_callables(obs.shape, meta["var"])
            if meta["kind"] == "admm":
                s = admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(),
                                          dimension=obs.ndim, alpha=meta["alpha"], rho=meta["rho"],
                                          iterations=meta["iterations"], iter_max=meta["iter_max"], x_scale=xs)
            elif meta["kind"] == "tikhonov":
                ident = lambda x: x.flatten()
                B, B_adj = (D, D_adj) if meta["reg"] == "TK1" else (ident, ident)
                s = tk.TikhonovLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=B, B_adj=B_adj, x0=obs.flatten(),
                                            alpha=meta["alpha"], iter_max=meta["iter_max"], x_scale=xs)
            else:
                s = make_pd_deconv(obs, meta["var"], meta["reg"], meta["alpha"], meta["iterations"], meta["iter_max"], xs,
                                   meta["L2"])
            s.run()
            assert rel_max(s.get_x(), golden("lsmr", name)) < F64_LSMR_TOL, (path, name)
        # observer path: every ADMM iterate
        from nsol_b200.observer import Observer
        obs = golden("lsmr", "in/bw2d")
        A, A_adj, D, D_adj = deconv_callables(obs.shape, [1.5, 1.5])
        Ao, Ao_adj, Do, Do_adj = orc.deconvolution_operators(obs.shape, np.diag([1.5, 1.5]))
        s = admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(), dimension=2,
                                  alpha=0.01, rho=0.5, iterations=4, iter_max=6, x_scale=float(obs.max()))
        o = Observer()
        s.set_observer(o)
        s.run()
        _, its = orc.admm_tv(Ao, Ao_adj, Do, Do_adj, obs.reshape(-1), obs.reshape(-1), 2, alpha=0.01, rho=0.5, iterations=4,
                             iter_max=6, x_scale=float(obs.max()), keep_iterates=True)
        assert len(o.get_x_list()) == 5
        for a, b in zip(o.get_x_list(), its):
            assert rel_max(a, b) < F64_LSMR_TOL
    finally:
        ctx.set_tuning("lsmr_path", 0)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("shape,var", [((64, 48), [1.0, 1.0]), ((40, 36), [0.4, 3.5]), ((12, 16, 20), [1.0, 0.5, 1.7]),
                                        ((130,), 1.0), ((24, 8), [1.0, 1.0])])
def test_lsmr_vector_kernels_match_generic_kernels(shape, var, dtype):
    """The radius-specialised 128-bit kernels (csrc/lsmr_fastv.cuh; lsmr_path 1) against the generic row-mapped
    kernels (lsmr_path 3) and the oracle: anisotropic masks (different radius per axis, incl. radius 1 and 6),
    rows as short as the mask allows, ADMM (B = grad) and Tikhonov TK0 (B = identity)."""
    rng = np.random.RandomState(23)
    obs = rng.rand(*shape) * 200 + 10
    xs = float(obs.max())
    A, A_adj, D, D_adj = deconv_callables(shape, var)
    ctx = _lib.context()
    out = {}
    try:
        for path in (1, 3):
            ctx.set_tuning("lsmr_path", path)
            s = admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(), dimension=len(shape),
                                      alpha=0.02, rho=0.3, iterations=3, iter_max=7, x_scale=xs, dtype=dtype)
            s.run()
            ident = lambda x: x.flatten()
            t = tk.TikhonovLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=ident, B_adj=ident, x0=obs.flatten(), alpha=0.05,
                                        iter_max=8, x_scale=xs, dtype=dtype)
            t.run()
            out[path] = (s.get_x(), t.get_x())
    finally:
        ctx.set_tuning("lsmr_path", 0)
    tol = 1e-11 if dtype == "float64" else F32_TOL
    assert rel_max(out[1][0], out[3][0]) < tol and rel_max(out[1][1], out[3][1]) < tol
    if dtype == "float64":
        cov = var if len(shape) == 1 else np.diag(var)
        Ao, Ao_adj, Do, Do_adj = orc.deconvolution_operators(shape, cov)
        ref = orc.admm_tv(Ao, Ao_adj, Do, Do_adj, obs.reshape(-1), obs.reshape(-1), len(shape), alpha=0.02, rho=0.3, iterations=3,
                          iter_max=7, x_scale=xs)
        assert rel_max(out[1][0], ref) < F64_LSMR_TOL


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("shape,var", [((40, 36), 1.0), ((70, 132), 1.0), ((33, 260), 0.4), ((9, 8), 1.0), ((64, 48), 3.5)])
def test_lsmr_fused_2d_kernels_match_generic_kernels(shape, var, dtype):
    """The fused 2-D forward / adjoint kernels (both blur passes inside the consumer: shared-memory x-blur +
    register ring along the rows; used for large images, forced here with lsmr_fuse2d = 1) against the generic
    kernels and the oracle: partial warp strips, several row chunks with periodic ring warm-up, radius 2, 3 and 6,
    the smallest image the ring allows."""
    rng = np.random.RandomState(29)
    obs = rng.rand(*shape) * 200 + 10
    xs = float(obs.max())
    A, A_adj, D, D_adj = deconv_callables(shape, [var, var])
    ctx = _lib.context()
    out = {}
    try:
        for tag, path, fuse in (("fused", 1, 1), ("fused_gen1", 1, 3), ("generic", 3, 2)):
            ctx.set_tuning("lsmr_path", path)
            ctx.set_tuning("lsmr_fuse2d", fuse)
            s = admm.ADMMLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(), dimension=2,
                                      alpha=0.02, rho=0.3, iterations=3, iter_max=7, x_scale=xs, dtype=dtype)
            s.run()
            tk1 = tk.TikhonovLinearSolver(A=A, A_adj=A_adj, b=obs.flatten(), B=D, B_adj=D_adj, x0=obs.flatten(), alpha=0.05,
                                          iter_max=8, x_scale=xs, dtype=dtype)
            tk1.run()
            out[tag] = (s.get_x(), tk1.get_x())
    finally:
        ctx.set_tuning("lsmr_path", 0)
        ctx.set_tuning("lsmr_fuse2d", 0)
    tol = 1e-11 if dtype == "float64" else F32_TOL
    assert rel_max(out["fused"][0], out["generic"][0]) < tol, rel_max(out["fused"][0], out["generic"][0])
    assert rel_max(out["fused"][1], out["generic"][1]) < tol
    assert rel_max(out["fused_gen1"][0], out["generic"][0]) < tol and rel_max(out["fused_gen1"][1], out["generic"][1]) < tol
    if dtype == "float64":
        Ao, Ao_adj, Do, Do_adj = orc.deconvolution_operators(shape, np.diag([var, var]))
        ref = orc.admm_tv(Ao, Ao_adj, Do, Do_adj, obs.reshape(-1), obs.reshape(-1), 2, alpha=0.02, rho=0.3, iterations=3,
                          iter_max=7, x_scale=xs)
        assert rel_max(out["fused"][0], ref) < F64_LSMR_TOL


# ------------------------------------------------------------------ measures on the device
def test_similarity_and_prior_measures_on_device():
    """SimilarityMeasures / PriorMeasures evaluated by device reductions equal the reference formulas
    (nsol/similarity_measures.py:26-120, nsol/prior_measures.py:19-52)."""
    from nsol_b200.similarity_measures import SimilarityMeasures as sim
    from nsol_b200.prior_measures import PriorMeasures as prior
    rng = np.random.RandomState(9)
    x_ref = rng.rand(37 * 29) * 255
    x = x_ref + rng.randn(x_ref.size) * 5
    n = float(x.size)
    ssd = np.sum(np.square(x - x_ref))
    expect = {"SSD": ssd, "MAE": np.sum(np.abs(x - x_ref)) / n, "MSE": ssd / n, "RMSE": np.sqrt(ssd / n),
              "PSNR": orc.psnr(x, x_ref), "NCC": orc.ncc(x, x_ref)}
    for name, val in expect.items():
        got = sim.similarity_measures[name](x, x_ref)
        assert abs(got - val) <= 1e-10 * abs(val), (name, got, val)
    assert abs(sim.sum_of_absolute_differences(x, x_ref) - np.sum(np.abs(x - x_ref))) < 1e-7
    # identities of tests/similarity_measures_test.py:31-93
    assert sim.sum_of_squared_differences(x, x) == 0 and np.isinf(sim.peak_signal_to_noise_ratio(x, x))
    assert abs(sim.normalized_cross_correlation(3 * x + 7, x) - sim.normalized_cross_correlation(x, x)) < 1e-12
    with pytest.raises(ValueError):
        sim.mean_squared_error(x, x_ref[:-1])
    # SSIM / MI / NMI are evaluated on the host (evaluation-only measures, nsol/similarity_measures.py:135-239)
    assert abs(sim.structural_similarity(x, x_ref) - orc.ssim_1d(x, x_ref)) < 1e-12
    assert sim.mutual_information(x, x_ref) > 0 and sim.normalized_mutual_information(x, x_ref) > 1
    shape = (37, 29)
    grad, _ = lo.LinearOperators2D(spacing=np.array([0.7, 1.3])).get_gradient_operators()
    D = lambda v: grad(v.reshape(*shape)).flatten()
    g = orc.grad(x.reshape(shape), [0.7, 1.3])
    ss = g[:37] ** 2 + g[37:] ** 2
    assert abs(prior.total_variation(x, D, 2) - np.sum(np.sqrt(ss))) < 1e-9 * np.sum(np.sqrt(ss))
    assert abs(prior.first_order_tikhonov(x, D) - 0.5 * np.sum(ss)) < 1e-9 * np.sum(ss)
    assert abs(prior.zeroth_order_tikhonov(x) - 0.5 * np.sum(x ** 2)) < 1e-9 * np.sum(x ** 2)
    hub = np.where(ss < 0.05 ** 2, ss, 2 * 0.05 * np.sqrt(ss) - 0.05 ** 2) / (2 * 0.05)
    assert abs(prior.huber(x, D, 2) - np.sum(hub)) < 1e-9 * np.sum(hub)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_observer_device_measures_match_host_measures(golden, dtype):
    """Observer(store_iterates=False): measures per iteration as device reductions, no iterate copies;
    same values as the reference flow (every iterate copied, measures on the host)."""
    from nsol_b200.observer import Observer
    from nsol_b200.similarity_measures import SimilarityMeasures as sim
    obs = golden("pd", "in/bw2d")
    x_ref = obs.reshape(-1) * 0.97 + 1.0
    names = ["PSNR", "RMSE", "NCC", "SSD"]
    out = {}
    for tag, store in (("host", True), ("device", False)):
        s = make_pd(obs, reg="TV", data="L2", alpha=0.05, L2=8, iterations=7, dtype=dtype)
        o = Observer(store_iterates=store)
        o.set_measures({m: (lambda x, m=m: sim.similarity_measures[m](x, x_ref)) for m in names})
        s.set_observer(o)
        s.run()
        o.compute_measures()
        out[tag] = (o.get_measures(), o.get_x_list())
    assert len(out["host"][1]) == 8 and len(out["device"][1]) == 1
    assert np.array_equal(out["device"][1][-1], out["host"][1][-1])
    for m in names:
        a, b = out["device"][0][m], out["host"][0][m]
        assert a.shape == (8,) and np.allclose(a, b, rtol=1e-9, atol=0), (m, a, b)
