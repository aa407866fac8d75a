"""Round-2 CPU tests (no GPU): the staged reference, host-side measures and losses, the deconvolution
study interface's wiring (symbolic probing only), the study writer against the files the reference's own
writer produced, bench.py's input generator and reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import nsol_oracle as orc
from oracle import ref_runner

import nsol_b200.admm_linear_solver as admm
import nsol_b200.linear_operators as lo
import nsol_b200.primal_dual_solver as pd
import nsol_b200.tikhonov_linear_solver as tk
from nsol_b200.deconvolution_solver_parameter_study_interface import (DeconvolutionParameterStudyInterface,
                                                                      DeconvolutionSolverStudyInterface)
from nsol_b200.loss_functions import LossFunctions
from nsol_b200.similarity_measures import SimilarityMeasures as sm


def test_staged_reference_is_the_unmodified_reference():
    """oracle/_ref (what the GPU box times as the CPU arm) is a byte-for-byte copy of /root/reference/nsol."""
    src = "/root/reference/nsol"
    dst = os.path.join(ROOT, "oracle", "_ref", "nsol")
    if not os.path.isdir(src) or not os.path.isdir(dst):
        pytest.skip("needs both /root/reference and the staged copy (build container)")
    import filecmp
    n = 0
    for root, _, files in os.walk(src):
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(root, f), src)
                assert filecmp.cmp(os.path.join(root, f), os.path.join(dst, rel), shallow=False), rel
                n += 1
    assert n >= 30


@pytest.mark.skipif(not ref_runner.available(), reason="reference not staged")
def test_reference_runner_matches_oracle_bit_for_bit():
    rng = np.random.RandomState(0)
    obs = rng.rand(9, 11, 10) * 255
    for reg, data in (("TV", "L2"), ("HUBER", "L1"), ("TK1", "L2")):
        s = ref_runner.pd_solver(obs, reg=reg, data=data, alpha=0.07, L2=8, iterations=7)
        dt, x = ref_runner.timed_run(s)
        ref = orc.primal_dual_denoise(obs.reshape(-1), obs.shape, reg=reg, data=data, alpha=0.07, L2=8, iterations=7,
                                      x_scale=float(obs.max()))
        assert np.array_equal(x, ref) and dt > 0


def test_r2_fixture_inputs_regenerate_bit_identically():
    """the full-size tests rebuild their inputs with the oracle's restatement of nsol/noise.py; the generator asserted
    equality with the reference's Noise class and recorded the blur-input deviation."""
    man = json.load(open(os.path.join(GOLDEN, "r2_manifest.json")))
    assert man["c3_input_oracle_vs_reference"] <= 1e-13
    for k in ("c2_full_alpha0.6", "c2_full_alpha0.05", "c4_128cube_100it", "c3_full_50x10", "c5_TV_a00", "c5_HUBER_a63", "c5_TK1_a27"):
        assert k in man and len(man[k]["sha256"]) == 64
    r2 = np.load(os.path.join(GOLDEN, "r2.npz"))
    z = np.load(os.path.join(GOLDEN, "inputs.npz"))
    # ADMM with b_reg: the oracle restatement reproduces the reference's result
    obs, c = r2["in/admm_breg_obs2"], r2["in/admm_breg_c2"]
    A, Aa, D, Da = orc.deconvolution_operators(obs.shape, np.eye(2))
    m = man["admm_breg_2d"]
    x = orc.admm_tv(A, Aa, D, Da, obs.reshape(-1), obs.reshape(-1), 2, alpha=m["alpha"], rho=m["rho"], iterations=m["iterations"],
                    iter_max=m["iter_max"], x_scale=float(obs.max()), b_reg=c)
    assert np.max(np.abs(x - r2["admm_breg_2d"])) <= 1e-12 * np.max(np.abs(x))
    x = orc.admm_tv(A, Aa, D, Da, obs.reshape(-1), obs.reshape(-1), 2, alpha=m["alpha"], rho=m["rho"], iterations=m["iterations"],
                    iter_max=m["iter_max"], x_scale=float(obs.max()), b_reg=1.5)
    assert np.max(np.abs(x - r2["admm_breg_2d_scalar"])) <= 1e-12 * np.max(np.abs(x))
    assert z["man_1024"].shape == (1024, 1024)


def test_oracle_matches_full_size_config2_hash():
    """the numpy restatement reproduces the reference's 1024^2 Huber-L1 result (200 iterations) bit for bit."""
    import hashlib
    man = json.load(open(os.path.join(GOLDEN, "r2_manifest.json")))
    z = np.load(os.path.join(GOLDEN, "inputs.npz"))
    sp = orc.add_salt_and_pepper_noise(z["man_1024"].astype(np.float64), 0.5, 0.1, seed=1)
    x = orc.primal_dual_denoise(sp.reshape(-1), sp.shape, reg="HUBER", data="L1", alpha=0.6, L2=8, iterations=200, x_scale=float(sp.max()))
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == man["c2_full_alpha0.6"]["sha256"]


# ------------------------------------------------------------------ host-side measures / losses
def test_host_measures():
    rng = np.random.RandomState(1)
    r = rng.rand(500) * 255
    a = r + rng.randn(500) * 3
    assert abs(sm.structural_similarity(a, r) - orc.ssim_1d(a, r)) < 1e-12
    assert abs(sm.structural_similarity(r, r) - 1.0) < 1e-12
    h = sm.shannon_entropy(r, bins=50)
    assert 0 < h <= np.log(50) + 1e-12
    assert abs(sm.mutual_information(r, r) - sm.shannon_entropy(r)) < 0.7     # MI(X, X) = H(X) up to the 2-D binning
    assert sm.normalized_mutual_information(a, r) > 1.0
    assert sm.mutual_information(a, r) > sm.mutual_information(rng.rand(500), r)
    x, y = r > 100, a > 100
    assert abs(sm.dice_score(x, y) - 2.0 * np.sum(x & y) / (x.sum() + y.sum())) < 1e-15
    assert sm.dice_score(x, x) == 1.0
    with pytest.raises(ValueError):
        sm.dice_score(r, a)
    with pytest.raises(ValueError):
        sm.structural_similarity(a, r[:-1])
    assert set(sm.similarity_measures) == {"SSD", "MAE", "MSE", "RMSE", "PSNR", "SSIM", "NCC", "MI", "NMI"}


def test_loss_functions_closed_forms():
    """tests/loss_functions_test.py:33-125 of the reference: each rho against its closed form."""
    f2 = np.linspace(0, 9, 37)
    for scale in (1.0, 2.5):
        z = f2 / scale ** 2
        assert np.allclose(LossFunctions.linear(f2, scale), f2)
        assert np.allclose(LossFunctions.soft_l1(f2, scale), scale ** 2 * 2 * (np.sqrt(1 + z) - 1))
        assert np.allclose(LossFunctions.huber(f2, scale), scale ** 2 * np.where(z <= 1, z, 2 * np.sqrt(z) - 1))
        assert np.allclose(LossFunctions.cauchy(f2, scale), scale ** 2 * np.log(1 + z))
        assert np.allclose(LossFunctions.arctan(f2, scale), scale ** 2 * np.arctan(z))
    f = np.array([1.0, -2.0, 3.0])
    assert LossFunctions.get_ell2_cost_from_residual(f) == 7.0
    assert set(LossFunctions.get_loss) == {"linear", "soft_l1", "huber", "cauchy", "arctan"}


# ------------------------------------------------------------------ deconvolution interface wiring (no GPU: probing only)
def _ops(shape):
    dim = len(shape)
    ops = getattr(lo, "LinearOperators%dD" % dim)()
    A, A_adj = ops.get_gaussian_blurring_operators(np.eye(dim) if dim > 1 else 1.0)
    g, g_adj = ops.get_gradient_operators()
    zshape = (dim * shape[0],) + tuple(shape[1:])
    w = lambda op, sh: (lambda x: op(x.reshape(*sh)).flatten())
    return w(A, shape), w(A_adj, shape), w(g, shape), w(g_adj, zshape)


@pytest.mark.parametrize("rtype,tv_solver,cls", [("TK0L2", "PD", tk.TikhonovLinearSolver), ("TK1L2", "PD", tk.TikhonovLinearSolver),
                                                 ("TVL2", "PD", pd.PrimalDualSolver), ("TVL2", "ADMM", admm.ADMMLinearSolver),
                                                 ("HuberL2", "PD", pd.PrimalDualSolver)])
def test_deconvolution_interface_builds_the_reference_wiring(rtype, tv_solver, cls):
    shape = (12, 14)
    A, A_adj, D, D_adj = _ops(shape)
    b = np.random.RandomState(0).rand(int(np.prod(shape))) * 100
    itf = DeconvolutionSolverStudyInterface(A=A, A_adj=A_adj, D=D, D_adj=D_adj, b=b, x0=b, alpha=0.02, x_scale=100.0, iter_max=7,
                                            iterations=9, minimizer="lsmr", measures=[], reconstruction_type=rtype, dimension=2,
                                            L2=8, rho=0.3, tv_solver=tv_solver)
    with pytest.raises(RuntimeError):
        itf.get_solver()
    with pytest.raises(RuntimeError):
        itf.get_measures()
    itf.set_up_solver()
    s = itf.get_solver()
    assert isinstance(s, cls) and s.get_alpha() == 0.02 and s.get_x_scale() == 100.0
    if cls is pd.PrimalDualSolver:
        cfg = s._probe()
        assert cfg["kind"] == "deconv" and cfg["reg"] == ("TV" if rtype == "TVL2" else "HUBER") and cfg["shape"] == shape
        assert cfg["lls"][7] == 7 and s.get_iterations() == 9 and s.get_L2() == 8.0
    elif cls is admm.ADMMLinearSolver:
        info = s._probe_lsq(s._B, s._B_adj)
        assert info["b_kind"] == "grad" and info["a_kind"] == "conv" and s.get_rho() == 0.3 and s.get_iterations() == 9
    else:
        info = s._probe_lsq(s._B, s._B_adj)
        assert info["b_kind"] == ("identity" if rtype == "TK0L2" else "grad") and s.get_iter_max() == 7
    itf.set_up_measures()
    assert sorted(itf.get_measures()) == ["Data", "Reg"]


def test_deconvolution_interface_errors_and_study_classes(tmp_path):
    shape = (8, 10)
    A, A_adj, D, D_adj = _ops(shape)
    b = np.ones(80)
    kw = dict(A=A, A_adj=A_adj, D=D, D_adj=D_adj, b=b, x0=b, alpha=0.02, x_scale=1.0, iter_max=5, iterations=3, minimizer="lsmr",
              measures=["PSNR"], dimension=2)
    with pytest.raises(KeyError):
        DeconvolutionSolverStudyInterface(reconstruction_type="TV", **kw).set_up_solver()
    with pytest.raises(ValueError):
        DeconvolutionSolverStudyInterface(reconstruction_type="TVL2", x_ref=[1, 2], **kw).set_up_measures()
    with pytest.raises(ValueError):
        DeconvolutionSolverStudyInterface(reconstruction_type="TVL2", x_ref=np.ones(79), **kw).set_up_measures()
    with pytest.raises(ValueError):
        DeconvolutionSolverStudyInterface(reconstruction_type="TVL2", x_ref=np.ones(80), x_ref_mask=np.ones(3), **kw).set_up_measures()
    from nsol_b200.admm_linear_solver_parameter_study import ADMMLinearSolverParameterStudy
    from nsol_b200.primal_dual_solver_parameter_study import PrimalDualSolverParameterStudy
    from nsol_b200.tikhonov_linear_solver_parameter_study import TikhonovLinearSolverParameterStudy
    for rtype, tv, cls in (("TK0L2", "PD", TikhonovLinearSolverParameterStudy), ("TK1L2", "PD", TikhonovLinearSolverParameterStudy),
                           ("TVL2", "PD", PrimalDualSolverParameterStudy), ("TVL2", "ADMM", ADMMLinearSolverParameterStudy),
                           ("HuberL2", "PD", PrimalDualSolverParameterStudy)):
        itf = DeconvolutionParameterStudyInterface(reconstruction_type=rtype, tv_solver=tv, dir_output=str(tmp_path), parameters={"alpha": [0.1]},
                                                   name=rtype, reconstruction_info={}, x_ref=np.ones(80), **kw)
        assert itf.get_parameter_study() is None
        itf.set_up_parameter_study()
        assert isinstance(itf.get_parameter_study(), cls)
        assert sorted(itf.get_measures()) == ["Data", "PSNR", "Reg"]


# ------------------------------------------------------------------ study writer vs the reference's own writer (CPU)
class _OracleSolver(object):
    """Stands in for the GPU solver on a box without one: same parameter-study protocol, iterates from the oracle."""

    def __init__(self, obs):
        self._obs, self._alpha, self._alg, self._observer = obs, 0.05, "ALG2", None

    def set_alpha(self, v):
        self._alpha = v

    def get_alpha(self):
        return self._alpha

    def set_alg_type(self, v):
        self._alg = v

    def get_alg_type(self):
        return self._alg

    def get_iterations(self):
        return 5

    def get_x_scale(self):
        return float(np.max(self._obs))

    def get_L2(self):
        return 8.0

    def set_observer(self, o):
        self._observer = o

    def get_x0(self):
        return self._obs.flatten()

    def set_x0(self, x0):
        pass

    def run(self):
        import datetime
        _, its = orc.primal_dual_denoise(self._obs.reshape(-1), self._obs.shape, reg="TV", data="L2", alpha=self._alpha, L2=8.0, iterations=5,
                                         x_scale=float(np.max(self._obs)), alg_type=self._alg, keep_iterates=True)
        for x in its:
            self._observer.add_x(x)
        self._observer.set_computational_time(datetime.timedelta(seconds=0.25))


def test_study_writer_matches_reference_writer_files(tmp_path):
    """SURVEY 8f row 4, pinned: tests/golden/study_ref was written by the reference's SolverParameterStudy
    (nsol/solver_parameter_study.py:229-323); this repo's writer, fed the same iterates, writes the same bytes
    (header time stamp and measured run times excepted) and the same reconstructions archive."""
    from nsol_b200.observer import Observer
    from nsol_b200.solver_parameter_study import SolverParameterStudy
    ref_dir = os.path.join(GOLDEN, "study_ref")
    inp = np.load(os.path.join(ref_dir, "input.npz"))
    noisy, clean = inp["noisy"], inp["clean"].flatten()

    class Study(SolverParameterStudy):
        def _get_fileheader(self):
            return self._header_from_keys(["alpha", "iterations", "x_scale", "L2"])     # PrimalDualSolverParameterStudy's keys
    obs = Observer()
    obs.set_measures({"SSD": lambda x: float(np.sum(np.square(x - clean))), "MAXABS": lambda x: float(np.max(np.abs(x)))})
    Study(solver=_OracleSolver(noisy), parameters={"alpha": [0.01, 0.05, 0.2], "alg_type": ["ALG2", "ALG3"]}, observer=obs,
          dir_output=str(tmp_path), name="RefStudy", reconstruction_info={"shape": noisy.shape}, append=False).run()
    strip = lambda l: l[:l.rindex("(")] if l.startswith("## ") and "(" in l else l
    for fname in ("RefStudy_parameters.txt", "RefStudy_measure_SSD.txt", "RefStudy_measure_MAXABS.txt"):
        ours = [strip(l) for l in open(os.path.join(str(tmp_path), fname)).read().split("\n")]
        theirs = [strip(l) for l in open(os.path.join(ref_dir, fname)).read().split("\n")]
        assert ours == theirs, fname
    ours = open(os.path.join(str(tmp_path), "RefStudy_computational_time.txt")).read().split("\n")
    theirs = open(os.path.join(ref_dir, "RefStudy_computational_time.txt")).read().split("\n")
    assert len(ours) == len(theirs) and [strip(l) for l in ours[:2]] == [strip(l) for l in theirs[:2]] and ours[2] == "0:00:00.250000"
    # header time stamp: two space-separated tokens, as the append check of the reference assumes (:117-118)
    assert len(open(os.path.join(ref_dir, "RefStudy_parameters.txt")).readline().split(" ")) == \
        len(open(os.path.join(str(tmp_path), "RefStudy_parameters.txt")).readline().split(" "))
    a = np.load(os.path.join(str(tmp_path), "RefStudy_reconstructions.npz"))
    b = np.load(os.path.join(ref_dir, "RefStudy_reconstructions.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in b.files:
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
    # and the reference-written study is readable by this repo's reader
    from nsol_b200.reader_parameter_study import ReaderParameterStudy
    reader = ReaderParameterStudy(ref_dir, "RefStudy")
    reader.read_study()
    assert reader.get_parameters() == {"alpha": [0.01, 0.05, 0.2], "alg_type": ["ALG2", "ALG3"]}
    assert reader.get_results("SSD").shape == (6, 6)


# ------------------------------------------------------------------ bench.py plumbing
def test_bench_input_is_slab_consistent_and_reference_arm_runs():
    sys.path.insert(0, ROOT)
    import bench
    whole = bench.synth_volume((16, 24, 24))
    part = bench.synth_volume((6, 24, 24), z_lo=7, nz_global=16)
    assert np.array_equal(whole[7:13], part)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-size", "48", "--ref-budget", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().split("\n")[-1])
    assert line["impl"] == "reference" and line["unit"] == "voxel-iterations/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_runner.available() else "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
    # a rank other than 0 prints nothing and exits 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_x0_is_observation_detection_without_gpu():
    rng = np.random.RandomState(4)
    obs = rng.rand(6, 8, 10) * 50
    from nsol_b200.proximal_operators import ProximalOperators as prox
    grad, grad_adj = lo.LinearOperators3D().get_gradient_operators()
    b = obs.flatten()
    mk = lambda x0: pd.PrimalDualSolver(prox_f=lambda x, tau: prox.prox_ell2_denoising(x, tau, x0=b, x_scale=50.0), prox_g_conj=prox.prox_tv_conj,
                                        B=lambda x: grad(x.reshape(6, 8, 10)).flatten(), B_conj=lambda x: grad_adj(x.reshape(18, 8, 10)).flatten(),
                                        L2=8, x0=x0, alpha=0.05, iterations=3, x_scale=50.0)
    s = mk(b)
    assert s._x0_is_observation(s._probe(), b)
    s2 = mk(b.copy())
    assert not s2._x0_is_observation(s2._probe(), b)
    s.set_x0(s.get_x0())          # what a parameter study does after every point: a fresh array
    assert not s._x0_is_observation(s._probe(), b)
    s3 = pd.PrimalDualSolver(prox_f=lambda x, tau: prox.prox_ell2_denoising(x, tau, x0=b, x_scale=25.0), prox_g_conj=prox.prox_tv_conj,
                             B=lambda x: grad(x.reshape(6, 8, 10)).flatten(), B_conj=lambda x: grad_adj(x.reshape(18, 8, 10)).flatten(),
                             L2=8, x0=b, alpha=0.05, iterations=3, x_scale=50.0)
    assert not s3._x0_is_observation(s3._probe(), b)      # different scales
