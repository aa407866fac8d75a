#!/usr/bin/env python
"""Round-2 fixtures from the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only).

    python oracle/gen_golden_r2.py [--only study|small|full]

Writes tests/golden/r2.npz + tests/golden/r2_manifest.json + tests/golden/study_ref/:
  * ADMMLinearSolver with b_reg != 0 (nsol/admm_linear_solver.py:100,171,208,222);
  * BASELINE configurations at their FULL size: C2 (1024^2 Huber-L1, 200 iterations, alpha 0.6 and 0.05),
    C3 (512^2 ADMM TV-L2, 50 x 10), C4 (128^3 TV-L2, 100 iterations), C5 (8 of the 64 alpha per regulariser
    TV / Huber / TK1 on 1024^2, 200 iterations).  Float64 results of bit-exact paths are stored as a SHA-256
    of their bytes plus a strided float64 subsample (a full 1024^2 array is 8 MB); C3 is stored in full;
  * a small parameter study written by the reference's own SolverParameterStudy file writers
    (nsol/solver_parameter_study.py:229-323) -- with import-only stand-ins for SimpleITK / natsort /
    matplotlib (oracle/import_stubs) and the pysitk file helpers of oracle/pysitk_stub.
Inputs are regenerated in the tests from tests/golden/inputs.npz with the oracle's restatement of
nsol/noise.py (bit-identical: same legacy RandomState stream), so only outputs are stored.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "import_stubs"))
from oracle import nsol_oracle as orc  # noqa: E402
from oracle import ref_runner as rr  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STRIDE = 97


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def ref_noise(data, kind, **kw):
    noise = rr.module("noise")
    n = noise.Noise(np.array(data, dtype=np.float64), seed=1)
    getattr(n, "add_%s_noise" % kind)(**kw)
    return n.get_noisy_data()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all")
    args = ap.parse_args()
    z = np.load(os.path.join(OUT, "inputs.npz"))
    path = os.path.join(OUT, "r2.npz")
    mpath = os.path.join(OUT, "r2_manifest.json")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    man = json.load(open(mpath)) if os.path.exists(mpath) else {}

    def keep(name, x, full=False, **params):
        x = np.asarray(x, dtype=np.float64)
        man[name] = dict(params, sha256=digest(x), size=int(x.size), stride=STRIDE, full=bool(full))
        out[name] = x if full else x[::STRIDE].copy()
        print("%-28s %s %s" % (name, man[name]["sha256"][:16], params), flush=True)

    def save():
        np.savez_compressed(path, **out)
        with open(mpath, "w") as fh:
            json.dump(man, fh, indent=1, sort_keys=True)

    if args.only in ("all", "small"):
        # ---- ADMM with b_reg != 0 ---------------------------------------------------------------
        rng = np.random.RandomState(11)
        lena = z["lena_512"].astype(np.float64)
        clean2 = lena[200:248, 220:260]
        A, _, _, _ = rr.deconv_ops(clean2.shape, np.eye(2))
        obs2 = ref_noise(A(clean2.flatten()).reshape(clean2.shape), "gaussian", noise_level=0.05)
        breg2 = rng.randn(2 * obs2.size) * 4.0
        out["in/admm_breg_obs2"], out["in/admm_breg_c2"] = obs2, breg2
        s = rr.admm_solver(obs2, [1.0, 1.0], alpha=0.02, rho=0.3, iterations=6, iter_max=8, b_reg=breg2)
        s.run()
        keep("admm_breg_2d", s.get_x(), full=True, var=[1.0, 1.0], alpha=0.02, rho=0.3, iterations=6, iter_max=8)
        s = rr.admm_solver(obs2, [1.0, 1.0], alpha=0.02, rho=0.3, iterations=6, iter_max=8, b_reg=1.5)
        s.run()
        keep("admm_breg_2d_scalar", s.get_x(), full=True, var=[1.0, 1.0], alpha=0.02, rho=0.3, iterations=6, iter_max=8, b_reg=1.5)
        ph = z["shepp_logan_64"].astype(np.float64)[16:36, 20:38, 18:42]
        A3, _, _, _ = rr.deconv_ops(ph.shape, np.eye(3))
        obs3 = ref_noise(A3(ph.flatten()).reshape(ph.shape), "gaussian", noise_level=0.05)
        breg3 = rng.randn(3 * obs3.size) * 3.0
        out["in/admm_breg_obs3"], out["in/admm_breg_c3"] = obs3, breg3
        s = rr.admm_solver(obs3, [1.0, 1.0, 1.0], alpha=0.01, rho=0.1, iterations=4, iter_max=10, b_reg=breg3)
        s.run()
        keep("admm_breg_3d", s.get_x(), full=True, var=[1.0, 1.0, 1.0], alpha=0.01, rho=0.1, iterations=4, iter_max=10)
        save()

    if args.only in ("all", "study"):
        # ---- study files written by the reference's own writer ----------------------------------------
        import shutil
        pdparam = rr.module("primal_dual_solver_parameter_study")
        observer_mod = rr.module("observer")
        sdir = os.path.join(OUT, "study_ref")
        shutil.rmtree(sdir, ignore_errors=True)
        os.makedirs(sdir)
        img = z["brainweb"][60:92, 50:78].astype(np.float64)
        noisy = ref_noise(img, "gaussian", noise_level=0.05)
        solver = rr.pd_solver(noisy, reg="TV", data="L2", alpha=0.05, L2=8, iterations=5)
        obs = observer_mod.Observer()
        ref = img.flatten()
        obs.set_measures({"SSD": lambda x: float(np.sum(np.square(x - ref))), "MAXABS": lambda x: float(np.max(np.abs(x)))})
        study = pdparam.PrimalDualSolverParameterStudy(
            solver, obs, dir_output=sdir, name="RefStudy",
            parameters={"alpha": [0.01, 0.05, 0.2], "alg_type": ["ALG2", "ALG3"]}, reconstruction_info={"shape": noisy.shape})
        study.run()
        np.savez_compressed(os.path.join(sdir, "input.npz"), noisy=noisy, clean=img)
        print("study files:", sorted(os.listdir(sdir)), flush=True)

    if args.only in ("all", "full"):
        man1024 = z["man_1024"].astype(np.float64)
        # ---- C2 full size ----------------------------------------------------------------------------
        sp = ref_noise(man1024, "salt_and_pepper", salt_vs_pepper=0.5, amount=0.1)
        assert np.array_equal(sp, orc.add_salt_and_pepper_noise(man1024, 0.5, 0.1, seed=1)), "oracle noise differs from the reference"
        for alpha in (0.6, 0.05):
            t0 = time.time()
            s = rr.pd_solver(sp, reg="HUBER", data="L1", alpha=alpha, L2=8, iterations=200)
            s.run()
            keep("c2_full_alpha%g" % alpha, s.get_x(), reg="HUBER", data="L1", alpha=alpha, L2=8, iterations=200, seconds=round(time.time() - t0, 1))
            save()
        # ---- C4 at 128^3 x 100 -------------------------------------------------------------------------
        ph = z["shepp_logan_64"].astype(np.float64)
        vol = np.repeat(np.repeat(np.repeat(ph, 2, 0), 2, 1), 2, 2)
        voln = ref_noise(vol, "gaussian", noise_level=0.05)
        assert np.array_equal(voln, orc.add_gaussian_noise(vol, 0.05, seed=1)), "oracle noise differs from the reference"
        t0 = time.time()
        s = rr.pd_solver(voln, reg="TV", data="L2", alpha=0.05, L2=8, iterations=100)
        s.run()
        keep("c4_128cube_100it", s.get_x(), reg="TV", data="L2", alpha=0.05, L2=8, iterations=100, seconds=round(time.time() - t0, 1))
        save()
        # ---- C3 full size: 512^2 ADMM, 50 x 10 ----------------------------------------------------------
        lena = z["lena_512"].astype(np.float64)
        A, _, _, _ = rr.deconv_ops(lena.shape, np.eye(2))
        obs = ref_noise(A(lena.flatten()).reshape(lena.shape), "gaussian", noise_level=0.05)
        Ao, _, _, _ = orc.deconvolution_operators(lena.shape, np.eye(2))
        obs_o = orc.add_gaussian_noise(Ao(lena.reshape(-1)).reshape(lena.shape), 0.05, seed=1)
        man["c3_input_oracle_vs_reference"] = float(np.max(np.abs(obs - obs_o)) / np.max(np.abs(obs)))
        t0 = time.time()
        s = rr.admm_solver(obs, [1.0, 1.0], alpha=0.01, rho=0.1, iterations=50, iter_max=10)
        s.run()
        keep("c3_full_50x10", s.get_x(), full=True, var=[1.0, 1.0], alpha=0.01, rho=0.1, iterations=50, iter_max=10, seconds=round(time.time() - t0, 1))
        save()
        # ---- C5: 8 of the 64 alpha per regulariser ----------------------------------------------------
        noisy = ref_noise(man1024, "gaussian", noise_level=0.05)
        assert np.array_equal(noisy, orc.add_gaussian_noise(man1024, 0.05, seed=1))
        alphas = np.linspace(0.001, 0.05, 64)
        for reg in ("TV", "HUBER", "TK1"):
            for i in (0, 9, 18, 27, 36, 45, 54, 63):
                name = "c5_%s_a%02d" % (reg, i)
                if name in man:
                    continue
                s = rr.pd_solver(noisy, reg=reg, data="L2", alpha=float(alphas[i]), L2=8, iterations=200)
                s.run()
                keep(name, s.get_x(), reg=reg, data="L2", alpha=float(alphas[i]), alpha_index=i, L2=8, iterations=200)
                save()
    save()
    print("written", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
