#!/usr/bin/env python
"""Benchmark of the NSoL proximal-solver hot path on B200 (BASELINE.json metric):
voxel-iterations/s of the fused 3-D TV-L2 primal-dual (Chambolle-Pock) iteration on the 512^3 volume
of BASELINE config 4, on 1/2/4/8 GPUs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--dtype float64|float32]
                    [--size 512] [--iters 100] [--scaling strong|weak] [--halo auto|p2p|nccl] [--no-extras]

One "step" = one full solve: reset (x = xbar = b/x_scale, p = 0) + ``iters`` fused primal-dual iterations
(alpha=0.05, L2=8, ALG2).
  value       device-timed (CUDA events on the launching stream, max over ranks) with the observation
              resident in HBM;
  e2e         the same solve through the public API -- PrimalDualSolver.run() + get_x(); at N > 1 the
              same calls on a solver sharded with PrimalDualSolver.distribute() -- with pinned HOST
              buffers, host<->device copies inside the timed region, over all --steps;
  roofline    algorithmic bytes (5+2d words per voxel-iteration) / measured average
              launch duration of the iteration kernel vs MEASURED_PEAKS.json;
  cpu_baseline  (N = 1) the UNMODIFIED reference's own PrimalDualSolver.run() (oracle/_ref, staged by
              oracle/make_ref.py; `kind: "reference"`) on a bounded sample of the same input, timed on
              this box's host cores;
  other_configs (N = 1) BASELINE configs 1, 2, 3, 5 and the large-problem ADMM/LSMR roofline records.
Multi-GPU (torchrun, one rank per GPU): the default is STRONG scaling -- the one 512^3 volume of
config 4 is z-slab sharded over the N ranks, 3 halo planes per interior boundary and iteration exchanged
inside the iteration kernels over NVLink peer memory (--halo p2p; NCCL send/recv with --halo nccl).
Every multi-GPU line carries `parity` (the sharded 512^3 result gathered on rank 0 and compared bit for
bit with the unsharded solve of the same volume) and a `weak` sub-record (512^3 per GPU).
``--impl reference`` times the reference's CPU implementation alone on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALPHA, L2 = 0.05, 8.0
UNIT = "voxel-iterations/s"
METRIC = "voxel-iterations/s, 3D TV-L2 primal-dual (Chambolle-Pock ALG2), BASELINE config 4 (512^3)"
METRIC_WEAK = "voxel-iterations/s, 3D TV-L2 primal-dual (Chambolle-Pock ALG2), weak scaling (512^3 per GPU)"
WORKLOAD = "C4: 3D TV-L2 primal-dual denoising %dx%dx%d, alpha=%g, L2=%g, ALG2, %d iterations per step"
INPUT = "64^3 Shepp-Logan fixture repeated to size + Gaussian noise 0.05 (seeded)"


def load_phantom():
    z = np.load(os.path.join(ROOT, "tests", "golden", "inputs.npz"))
    return z["shepp_logan_64"].astype(np.float64)


def synth_volume(shape, z_lo=0, nz_global=None, seed=1):
    """BASELINE config 4 input: the reference's 64^3 Shepp-Logan fixture repeated along each axis to the
    global shape (nz_global, ny, nx) plus Gaussian noise 0.05 (nsol/noise.py:51-55); returns the planes
    [z_lo, z_lo + shape[0]).  The noise of plane z is seeded by (seed, z), so a z-slab of the volume is the
    same data whichever rank generates it."""
    ph = load_phantom()
    nz_global = shape[0] if nz_global is None else nz_global
    gshape = (nz_global, shape[1], shape[2])
    reps = [int(np.ceil(s / 64.0)) for s in gshape]
    zi = (np.arange(z_lo, z_lo + shape[0]) // reps[0]).clip(0, 63)
    yi = (np.arange(shape[1]) // reps[1]).clip(0, 63)
    xi = (np.arange(shape[2]) // reps[2]).clip(0, 63)
    vol = ph[np.ix_(zi, yi, xi)].astype(np.float64)
    for k in range(shape[0]):
        rng = np.random.RandomState([seed, z_lo + k])
        vol[k] += 0.05 * 255.0 * rng.standard_normal(size=vol.shape[1:])
    return vol


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons sampled every 50 ms during the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                out["samples"] = len(sm)
                out["power_w_max"] = max(float(r[3]) for r in rows if len(r) >= 9)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, nm in enumerate(names):
                    if any(r[5 + i].strip().lower() == "active" for r in rows if len(r) >= 9):
                        out["reasons"].append(nm)
        except Exception:
            pass
        return out


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel_key, nvox):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of
    this kernel at this chunking, committed under profiles/), scaled to the voxels of this launch."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                t = json.load(fh)[kernel_key]
            return {"bytes": t["bytes_per_launch"] * (nvox / float(t["voxels"])), "source": "profiles/" + t.get("source", name)}
        except Exception:
            continue
    return None


def device_copy_bandwidth(nbytes=1 << 31):
    """Device-to-device copy bandwidth (read + write bytes) of THIS GPU in this run, measured the way
    MEASURED_PEAKS.json was (torch b.copy_(a), best of 10)."""
    import torch
    a = torch.empty(nbytes // 2, dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    a.zero_()
    best = None
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    del a, b
    return 2.0 * nbytes / (best * 1e-3) / 1e9


def pcie_bandwidth(ctx, stream, nbytes=1 << 28):
    """Measured pinned host<->device copy bandwidth of this box (explains the e2e number)."""
    import torch
    host = ctx.pinned_empty((nbytes,), np.uint8)
    dev = ctx.device_alloc(nbytes)
    out = {}
    for name, fn in (("h2d_gbs", lambda: ctx.lib.nsol_memcpy_h2d(ctx.handle, dev.ptr, host.ctypes.data, nbytes, stream)),
                     ("d2h_gbs", lambda: ctx.lib.nsol_memcpy_d2h(ctx.handle, host.ctypes.data, dev.ptr, nbytes, stream))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = 3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    dev.free()
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref), else the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_reference_solver(sample_n, iters):
    """(kind, run) -- run() executes ONE solve of ``iters`` iterations on a sample_n^3 sub-volume of the
    config-4 input and returns the seconds solver.run() took.  kind "reference": the reference's own
    PrimalDualSolver wired as nsol/application/run_denoising.py:95-154 (nsol/primal_dual_solver.py:215-263);
    kind "port": the oracle's restatement with the reference's call structure (only if oracle/_ref is absent)."""
    vol = synth_volume((sample_n, sample_n, sample_n))
    try:
        from oracle import ref_runner
        have_ref = ref_runner.available()
    except Exception:
        have_ref = False
    if have_ref:
        def run():
            solver = ref_runner.pd_solver(vol, reg="TV", data="L2", alpha=ALPHA, L2=L2, iterations=iters)
            return ref_runner.timed_run(solver)[0]
        return "reference", run
    from oracle import nsol_oracle as orc

    def run():
        t0 = time.perf_counter()
        orc.primal_dual_denoise_ndimage(vol.reshape(-1), vol.shape, reg="TV", data="L2", alpha=ALPHA, L2=L2, iterations=iters,
                                        x_scale=float(vol.max()))
        return time.perf_counter() - t0
    return "port", run


def cpu_baseline_record(sample_n, budget_s, steps=1):
    """Time the CPU arm on a bounded sample: about ``budget_s`` seconds in total."""
    kind, run1 = cpu_reference_solver(sample_n, 1)
    t1 = run1()                                         # one iteration: calibrates the sample (and warms up)
    iters = int(max(1, min(4, budget_s / max(steps, 1) / max(t1, 1e-3))))
    _, run = cpu_reference_solver(sample_n, iters)
    times = [run() for _ in range(steps)]
    dt = float(np.mean(times))
    what = ("the unmodified reference (gift-surg/NSoL v0.1.14, oracle/_ref): PrimalDualSolver.run(), numpy + scipy.ndimage, "
            "single-threaded" if kind == "reference" else
            "numpy/scipy.ndimage port with the reference's call structure (oracle/_ref not staged), single-threaded")
    return {"value": sample_n ** 3 * iters / dt, "unit": UNIT, "cores": 1, "kind": kind, "host_cores": os.cpu_count(),
            "sample": "%d^3 sub-volume of the config-4 input, %d iteration(s) per step, %d step(s), %.1f s per step; %s"
                      % (sample_n, iters, steps, dt, what),
            "ms_per_step": dt * 1e3, "iterations": iters, "versions": {"numpy": np.__version__, "scipy": __import__("scipy").__version__}}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    rec = cpu_baseline_record(args.ref_size, args.ref_budget, steps=max(1, args.steps))
    n = args.size
    line = {
        "impl": "reference", "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % (n, n, n, ALPHA, L2, args.iters), "input": INPUT, "cpu_sample": rec["sample"]},
        "cpu_baseline": {k: rec[k] for k in ("value", "unit", "cores", "kind", "host_cores", "sample", "versions")},
        "e2e": {"value": rec["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# other BASELINE configurations and the ADMM / LSMR roofline records (N = 1)
# ------------------------------------------------------------------------------------------------
def time_events(fn, reps):
    import torch
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def other_configs(ctx, stream, peak):
    """Device-timed runs of BASELINE configurations 1, 2, 3, 5 on one GPU: latency-bound 2-D single images,
    the ADMM/LSMR deconvolution and the batched alpha sweep.  Inputs are the reference's test images
    (tests/golden/inputs.npz) with seeded synthetic noise."""
    from nsol_b200 import _lib
    import nsol_b200.linear_operators as lo
    import nsol_b200.admm_linear_solver as admm
    z = np.load(os.path.join(ROOT, "tests", "golden", "inputs.npz"))
    rng = np.random.RandomState(1)
    lib = ctx.lib
    out = {}

    def pd_case(img, reg, data, alphas, iters, dtype="float64"):
        dcode = _lib.dtype_code(dtype)
        esz = 4 if dcode == _lib.F32 else 8
        desc = _lib.PdDesc()
        desc.grid = _lib.make_grid(img.shape, None, dcode, len(alphas))
        desc.reg, desc.data, desc.alg = _lib.REG[reg], _lib.DATA[data], _lib.ALG["ALG2"]
        desc.huber_gamma, desc.L2 = 0.05, 8.0
        desc.x_scale = desc.x0_scale = desc.b_scale = float(img.max())
        al = np.ascontiguousarray(alphas, dtype=np.float64)
        desc.alpha = al.ctypes.data_as(_lib.c_double_p)
        plan = C.c_void_p()
        ctx.check(lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(plan)))
        host = np.ascontiguousarray(img.reshape(-1), dtype=np.float64)

        def solve():
            ctx.check(lib.nsol_pd_plan_reset_host(plan, host.ctypes.data, None, stream))
            ctx.check(lib.nsol_pd_plan_iterate(plan, iters, stream))
        ms = time_events(solve, 3)
        lib.nsol_pd_plan_destroy(plan)
        vox_it = img.size * len(alphas) * iters
        words = 5 + 2 * img.ndim
        return {"ms_per_solve": ms, "voxel_iters_per_s": vox_it / (ms * 1e-3), "iterations": iters, "batch": len(alphas),
                "us_per_iteration": ms * 1e3 / iters,
                "one_pass_hbm_model_frac": words * esz * vox_it / (ms * 1e-3) / 1e9 / peak, "dtype": dtype,
                "note": "includes the H2D copy of the observation and the reset; one_pass_hbm_model_frac = (5 + 2d words per "
                        "pixel-iteration) / time / measured HBM bandwidth -- 2-D solves run K iterations per pass over the state "
                        "(temporal blocking, csrc/pd_tb2d.cuh), so a value above 1 means fewer passes, not more bandwidth"}

    lena = z["lena_256_noise"].astype(np.float64)
    out["C1_2D_TVL2_PD_lena256_100it"] = pd_case(lena, "TV", "L2", [0.05], 100)
    man = z["man_1024"].astype(np.float64)
    sp = man.copy().reshape(-1)
    idx = rng.choice(sp.size, size=int(0.1 * sp.size), replace=False)
    sp[idx[:idx.size // 2]] = man.max()
    sp[idx[idx.size // 2:]] = man.min()
    out["C2_2D_HuberL1_PD_man1024_200it"] = pd_case(sp.reshape(man.shape), "HUBER", "L1", [0.6], 200)
    noisy = man + 0.05 * man.max() * rng.standard_normal(man.shape)
    alphas = np.linspace(0.001, 0.05, 64)
    for reg in ("TV", "HUBER", "TK1"):
        out["C5_sweep64_%sL2_PD_man1024_200it" % reg] = pd_case(noisy, reg, "L2", alphas, 200)

    # config 3: ADMM TV-L2 deconvolution, 512^2, sigma=1 periodic blur, 50 outer x 10 LSMR iterations
    img = z["lena_512"].astype(np.float64)
    ops = lo.LinearOperators2D()
    A, A_adj = ops.get_gaussian_blurring_operators(np.eye(2))
    grad, grad_adj = ops.get_gradient_operators()
    obs = A(img) + 0.05 * img.max() * rng.standard_normal(img.shape)
    shape = img.shape
    zshape = (2 * shape[0], shape[1])
    solver = admm.ADMMLinearSolver(
        A=lambda x: A(x.reshape(*shape)).flatten(), A_adj=lambda x: A_adj(x.reshape(*shape)).flatten(), b=obs.flatten(),
        B=lambda x: grad(x.reshape(*shape)).flatten(), B_adj=lambda x: grad_adj(x.reshape(*zshape)).flatten(),
        x0=obs.flatten(), dimension=2, alpha=0.01, rho=0.1, iterations=50, iter_max=10, x_scale=float(obs.max()))
    solver.run()
    t0 = time.perf_counter()
    for _ in range(3):
        solver.run()
        solver.get_x()
    dt = (time.perf_counter() - t0) / 3
    out["C3_2D_TVL2_ADMM_lena512_50x10"] = {
        "ms_per_solve": dt * 1e3, "pixel_lsmr_iters_per_s": img.size * 50 * 10 / dt, "us_per_inner_iteration": dt * 1e6 / 500,
        "note": "ADMMLinearSolver.run()+get_x() wall clock (host buffers; outer iterations replayed as a CUDA graph)"}
    solver.release()
    return out


def lsmr_roofline_records(ctx, stream, peak):
    """ADMM TV-L2 deconvolution on problems far larger than L2, device-resident: time per inner LSMR iteration
    against the 19-word (2-D) / 22-word (3-D) algorithmic traffic of SURVEY.md 8d (per-outer-iteration kernels
    -- right-hand side, start vectors, clip, shrink -- included in the time, not in the bytes)."""
    from nsol_b200 import _lib
    from nsol_b200.linear_solver import LsmrPlan
    import nsol_b200.kernels as kern
    lib = ctx.lib
    out = {}
    for name, shape, dtype in (("ADMM_2D_4096x4096_f64", (4096, 4096), "float64"), ("ADMM_3D_256x256x256_f64", (256, 256, 256), "float64"),
                               ("ADMM_3D_256x256x256_f32", (256, 256, 256), "float32")):
        dim = len(shape)
        dcode = _lib.dtype_code(dtype)
        esz = 4 if dcode == _lib.F32 else 8
        n = int(np.prod(shape))
        # the reference's Gaussian mask for cov = identity, spacing 1 (radius 3), as separable taps
        mask = getattr(kern, "Kernels%dD" % dim)().get_gaussian(np.eye(dim))

        class _Op(object):
            pass
        a_op = _Op()
        a_op.taps = kern.separable_taps(mask)
        info = {"shape": shape, "spacing": (1.0,) * dim, "a_kind": "conv", "a_op": a_op, "b_kind": "grad", "dim": dim}
        plan = LsmrPlan(info, dtype)
        rng = np.random.RandomState(7)
        host = (rng.rand(n) * 0.9 + 0.05).astype(_lib.np_dtype(dcode))
        b = ctx.device_alloc(n * esz).upload(host)
        x0 = ctx.device_alloc(n * esz).upload(host)
        x = ctx.device_alloc(n * esz)
        outer, inner = 2, 10

        def solve():
            ctx.check(lib.nsol_admm_run_dev(plan.handle, 0.01, 0.1, outer, inner, b.ptr, x0.ptr, x.ptr, stream))
        l0 = plan.ctx.launch_count()
        ms = time_events(solve, 2)
        launches = (plan.ctx.launch_count() - l0) // 3
        us = ms * 1e3 / (outer * inner)
        # 19 (2-D) / 22 (3-D) words per voxel and inner iteration, plus per OUTER iteration (1 + 3d) words for the
        # shrink / w update (K7) and ~12 words of LSMR start-up (right-hand side, first adjoint, start vectors, clip)
        words = 19 if dim == 2 else 22
        words_outer = (1 + 3 * dim) + 12
        words_all = words + words_outer / float(inner)
        out[name] = {"us_per_inner_iteration": us, "voxel_lsmr_iters_per_s": n / (us * 1e-6), "words_per_voxel": words,
                     "achieved_gbs": words * esz * n / (us * 1e-6) / 1e9, "hbm_frac_of_measured": words * esz * n / (us * 1e-6) / 1e9 / peak,
                     "words_per_voxel_incl_outer_iteration": words_all,
                     "hbm_frac_of_measured_incl_outer_iteration": words_all * esz * n / (us * 1e-6) / 1e9 / peak,
                     "launches_per_solve": launches, "outer_x_inner": "%dx%d" % (outer, inner), "dtype": dtype,
                     "note": "the time includes the per-outer-iteration kernels; 'words_per_voxel' counts the inner iteration only"}
        for buf in (b, x0, x):
            buf.free()
        plan.close()
    return out


def admm_slab_record(ctx, rank, world, device, peak):
    """N > 1: ADMM TV-L2 deconvolution of ONE 512x256x256 float64 volume, z-slab sharded over the ranks (blur halos on a
    ring, gradient halos between neighbours, all-reduced LSMR norms; the outer iteration replayed as a CUDA graph that
    contains the NCCL calls).  Device-timed (CUDA events, max over ranks) voxel-LSMR-iterations/s."""
    import torch
    import torch.distributed as dist
    import nsol_b200.kernels as kern
    from nsol_b200.distributed import SlabADMM, slab_bounds
    gshape = (512, 256, 256)
    z_lo, z_hi = slab_bounds(gshape[0], rank, world)
    shape = (z_hi - z_lo,) + gshape[1:]
    mask = kern.Kernels3D().get_gaussian(np.eye(3))

    class _Op(object):
        pass
    a_op = _Op()
    a_op.taps = kern.separable_taps(mask)
    info = {"shape": shape, "spacing": (1.0, 1.0, 1.0), "a_kind": "conv", "a_op": a_op, "b_kind": "grad", "dim": 3}
    rng = np.random.RandomState([5, rank])
    host = rng.rand(int(np.prod(shape))) * 0.9 + 0.05
    slab = SlabADMM(ctx, info, "float64", rank, world, device)
    outer, inner = 5, 10
    slab.run(host, host, 0.01, 0.1, outer, inner)            # warm-up (graph capture, NCCL channels)
    times = []
    for _ in range(2):
        torch.cuda.synchronize()
        dist.barrier()
        x = slab.run(host, host, 0.01, 0.1, outer, inner)
        t = torch.tensor([slab.last_device_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
    cs = torch.tensor([float(np.sum(x[::997]))], dtype=torch.float64, device=device)
    dist.all_reduce(cs)
    used_graph = slab.last_used_graph
    slab.close()
    ms = min(times)
    n = int(np.prod(gshape))
    us = ms * 1e3 / (outer * inner)
    return {"workload": "ADMM TV-L2 deconvolution %dx%dx%d float64, sigma=1 periodic blur, %d outer x %d LSMR iterations, z-slab x%d" % (gshape + (outer, inner, world)),
            "us_per_inner_iteration": us, "voxel_lsmr_iters_per_s": n / (us * 1e-6), "ms_per_solve": ms,
            "hbm_frac_of_measured_per_gpu": 22 * 8 * (n / world) / (us * 1e-6) / 1e9 / peak,
            "communication": "per inner iteration: 2 grouped NCCL send/recv rounds (ring blur halos, neighbour gradient halos) + 3 one-double "
                             "all-reduces, %s" % ("all inside one replayed CUDA graph per outer iteration" if used_graph else "issued from the host"),
            "checksum": float(cs.item())}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--iters", type=int, default=100, help="primal-dual iterations per step (solve)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the ONE 512^3 volume of BASELINE config 4 sharded over N GPUs (default, the "
                         "BASELINE metric); weak = 512^3 per GPU (always reported in the 'weak' sub-record)")
    ap.add_argument("--planes", type=int, default=0, help="experiments only: z-planes of the (strong-scaling) volume instead of --size")
    ap.add_argument("--ref-size", type=int, default=256)
    ap.add_argument("--ref-budget", type=float, default=100.0, help="seconds of CPU time the reference arm may spend in total")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs / lsmr_roofline (N=1) and weak / parity (N>1)")
    ap.add_argument("--halo", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU halo exchange: p2p = inside the iteration kernels over peer memory (NVLink), "
                         "nccl = grouped send/recv before every launch; auto = p2p if CUDA IPC works")
    ap.add_argument("--secondary-dtype", action="store_true", help="also time the other dtype (reported under 'other_dtype')")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from nsol_b200 import _lib
    from nsol_b200.distributed import SlabPrimalDual, slab_bounds

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    ctx = _lib.context(local_rank)
    n = args.size
    if os.environ.get("NSOL_BENCH_OWN_STREAM", "0") == "1":      # experiments: a created stream instead of the legacy default stream
        torch.cuda.set_stream(torch.cuda.Stream())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib = ctx.lib

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    halo_mode = ["single"]

    def make_desc(shape, dcode, x_scale):
        desc = _lib.PdDesc()
        desc.grid = _lib.make_grid(shape, None, dcode, 1)
        desc.reg, desc.data, desc.alg = _lib.REG["TV"], _lib.DATA["L2"], _lib.ALG["ALG2"]
        desc.b_batched = 0
        desc.huber_gamma, desc.L2 = 0.05, L2
        desc.x_scale = desc.x0_scale = desc.b_scale = x_scale
        alpha_arr = np.array([ALPHA])
        desc.alpha = alpha_arr.ctypes.data_as(_lib.c_double_p)
        desc._keep = alpha_arr
        return desc

    def measure(dtype_name, scaling, steps, want_e2e, want_parity):
        """One workload: global volume (nz_global, n, n) sharded over the ranks."""
        nz_global = n * world if scaling == "weak" else (args.planes if args.planes > 0 else n)
        z_lo, z_hi = slab_bounds(nz_global, rank, world)
        nz_loc = z_hi - z_lo
        shape = (nz_loc, n, n)
        nvox_loc = nz_loc * n * n
        nvox_global = nz_global * n * n
        dcode = _lib.dtype_code(dtype_name)
        esz = 4 if dcode == _lib.F32 else 8
        np_dt = _lib.np_dtype(dcode)
        # ---- synthetic observation, pinned on the host, then resident in HBM -------------
        obs_host = ctx.pinned_empty((nvox_loc,), np.float64)
        vol = synth_volume(shape, z_lo=z_lo, nz_global=nz_global)
        x_scale = max_over_ranks(float(vol.max()))   # x_scale = max(observed) (nsol/application/run_denoising.py:97)
        obs_host[:] = vol.reshape(-1)
        del vol
        desc = make_desc(shape, dcode, x_scale)
        slab = SlabPrimalDual(ctx, desc, n * n, np_dt, rank, world, device, halo=args.halo)
        halo_mode[0] = slab.mode
        plan = slab.plan
        stage = ctx.device_alloc(nvox_loc * 8)
        obs_dev = ctx.device_alloc(nvox_loc * esz)
        ctx.check(lib.nsol_memcpy_h2d(ctx.handle, stage.ptr, obs_host.ctypes.data, nvox_loc * 8, stream))
        ctx.check(lib.nsol_scale_convert(ctx.handle, nvox_loc, _lib.F64, stage.ptr, dcode, obs_dev.ptr, 1.0, 0, stream))
        torch.cuda.synchronize()
        stage.free()

        for _ in range(args.warmup):
            slab.reset_dev(obs_dev.ptr, None, stream, check=False)
            slab.iterate(args.iters, stream)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = ctx.launch_count()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        pairs = []
        for _ in range(steps):
            e1, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            slab.reset_dev(obs_dev.ptr, None, stream, check=False)
            e1.record()
            slab.iterate(args.iters, stream)
            e2.record()
            pairs.append((e1, e2))
        ev1.record()
        barrier()
        slab.check(stream)      # a timed-out in-kernel halo wait invalidates the run
        total_ms = max_over_ranks(ev0.elapsed_time(ev1))
        kernel_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in pairs))
        launches = ctx.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        res = {
            "value": nvox_global * args.iters * steps / (total_ms * 1e-3),
            "ms_per_step": total_ms / steps, "steps": steps,
            "iter_ms": kernel_ms / (steps * args.iters),
            "launches": launches, "clocks": clocks, "nvox_loc": nvox_loc, "nvox_global": nvox_global, "nz_global": nz_global,
            "plan_bytes": int(lib.nsol_pd_plan_bytes(plan)), "chunks": int(lib.nsol_pd_plan_chunks(plan)),
        }
        # ---- parity: sharded result gathered on rank 0 vs the unsharded solve of the same volume ------------
        if want_parity and world > 1:
            res["parity"] = slab_parity(slab, obs_dev, desc, dcode, esz, np_dt, nz_global, nz_loc, x_scale)
        slab.close()
        obs_dev.free()
        # ---- end to end through the public API, host buffers ---------------------------
        if want_e2e:
            import nsol_b200.linear_operators as lo
            import nsol_b200.primal_dual_solver as pd
            from nsol_b200.proximal_operators import ProximalOperators as prox
            grad, grad_adj = lo.LinearOperators3D().get_gradient_operators()
            zshape = (3 * shape[0],) + shape[1:]
            D = lambda x: grad(x.reshape(*shape)).flatten()
            D_adj = lambda x: grad_adj(x.reshape(*zshape)).flatten()
            b = obs_host
            solver = pd.PrimalDualSolver(
                prox_f=lambda x, tau: prox.prox_ell2_denoising(x, tau, x0=b, x_scale=x_scale),
                prox_g_conj=prox.prox_tv_conj, B=D, B_conj=D_adj, L2=L2, x0=b, alpha=ALPHA,
                iterations=args.iters, x_scale=x_scale, dtype=dtype_name)
            if world > 1:
                solver.distribute(halo=args.halo)

            def e2e_step():
                solver.run()
                return solver.get_x()
            # warm-up: two results alive at once, so the page-locked result pool holds a spare buffer
            # and no timed step pays for a cudaHostAlloc
            x = e2e_step()
            x_spare = e2e_step()
            del x_spare
            barrier()
            link = pcie_bandwidth(ctx, stream) if rank == 0 else None
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                x = e2e_step()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            same = solver._x0_is_observation(solver._probe(), np.asarray(b))
            res["e2e"] = {"value": nvox_global * args.iters * steps / dt, "unit": UNIT,
                          "h2d_bytes_per_step": (1 if same else 2) * nvox_loc * 8, "d2h_bytes_per_step": nvox_loc * 8,
                          "bytes_are": "per rank" if world > 1 else "total",
                          "ms_per_step": dt / steps * 1e3, "steps": steps,
                          "api": "PrimalDualSolver.run()+get_x()" + (" on a solver sharded with PrimalDualSolver.distribute()" if world > 1 else ""),
                          "host_link": link}
            cs = torch.tensor([float(np.sum(x[:: max(1, x.size // 4096)]))], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(cs)
            res["checksum"] = float(cs.item())
            del x
            solver.release()
        return res

    def slab_parity(slab, obs_dev, desc, dcode, esz, np_dt, nz_global, nz_loc, x_scale):
        """Run ONE more sharded solve, gather x on rank 0, and compare it bit for bit with the unsharded solve of
        the whole volume on rank 0's GPU."""
        nvox_loc = nz_loc * n * n
        slab.reset_dev(obs_dev.ptr, None, stream)
        slab.iterate(args.iters, stream)
        slab.check(stream)
        xptr = C.c_void_p()
        ctx.check(lib.nsol_pd_plan_x_dev(slab.plan, C.byref(xptr)))
        from nsol_b200.distributed import tensor_from_ptr
        x_loc = tensor_from_ptr(xptr.value, nvox_loc, np_dt, device)
        b_loc = tensor_from_ptr(obs_dev.ptr.value if hasattr(obs_dev.ptr, "value") else obs_dev.ptr, nvox_loc, np_dt, device)
        sizes = [(slab_bounds(nz_global, r, world)[1] - slab_bounds(nz_global, r, world)[0]) * n * n for r in range(world)]
        tdt = torch.float32 if np.dtype(np_dt) == np.float32 else torch.float64
        out = None
        gathered = {}
        for name, t in (("x", x_loc), ("b", b_loc)):
            if rank == 0:
                parts = [torch.empty(s, dtype=tdt, device=device) for s in sizes]
                parts[0].copy_(t)
                for r in range(1, world):
                    dist.recv(parts[r], src=r)
                gathered[name] = torch.cat(parts)
                del parts
            else:
                dist.send(t.clone(), dst=0)
        if rank == 0:
            gshape = (nz_global, n, n)
            gdesc = make_desc(gshape, dcode, x_scale)
            h = C.c_void_p()
            ctx.check(lib.nsol_pd_plan_create(ctx.handle, C.byref(gdesc), C.byref(h)))
            ctx.check(lib.nsol_pd_plan_reset_dev(h, C.c_void_p(gathered["b"].data_ptr()), None, stream))
            ctx.check(lib.nsol_pd_plan_iterate(h, args.iters, stream))
            ctx.check(lib.nsol_pd_plan_x_dev(h, C.byref(xptr)))
            torch.cuda.synchronize()
            x_ref = tensor_from_ptr(xptr.value, nz_global * n * n, np_dt, device)
            diff = float((gathered["x"] - x_ref).abs().max().item())
            equal = bool(torch.equal(gathered["x"], x_ref))
            out = {"check": "z-slab sharded solve (%d ranks, halo %s) vs the unsharded solve of the same %dx%dx%d volume on rank 0, %d iterations"
                            % (world, slab.mode, nz_global, n, n, args.iters),
                   "bit_identical": equal, "max_abs_diff": diff, "x_abs_max": float(x_ref.abs().max().item())}
            lib.nsol_pd_plan_destroy(h)
        del gathered
        barrier()
        return out

    # ---- measurements ------------------------------------------------------------------------------
    head_scaling = args.scaling          # at N = 1 both are the same 512^3 volume
    main_res = measure(args.dtype, head_scaling, args.steps, not args.no_e2e, not args.no_extras)
    weak_res = None
    if world > 1 and not args.no_extras:
        other_scaling = "weak" if head_scaling == "strong" else "strong"
        weak_res = measure(args.dtype, other_scaling, max(2, min(args.steps, 5)), not args.no_e2e, False)
    other = None
    if args.secondary_dtype:
        od = "float32" if args.dtype == "float64" else "float64"
        other = measure(od, head_scaling, args.steps, False, False)

    peak, peak_src = measured_peak()
    admm_rec = None
    if world > 1 and not args.no_extras:
        try:
            admm_rec = admm_slab_record(ctx, rank, world, device, peak)
        except Exception as e:      # never lose the headline line
            admm_rec = {"error": "%s: %s" % (type(e).__name__, e)}
    copy_gbs = device_copy_bandwidth() if rank == 0 else None
    if rank == 0:
        esz = 4 if args.dtype == "float32" else 8
        nvox_loc = main_res["nvox_loc"]
        achieved = 11 * esz * nvox_loc / (main_res["iter_ms"] * 1e-3) / 1e9
        # default variants (csrc/pd_kernels.cu): TMA bulk-async kernel for float64 3-D, register-pipelined LDG for float32
        variant = int(os.environ.get("NSOL_PD_VARIANT", "0")) or (2 if args.dtype == "float64" else 1)
        link = 1 if halo_mode[0] == "p2p" else 0
        kernel_name = ("pd_iter_bulk_kernel" if variant == 2 else "pd_iter_kernel") + \
            "<%s, TV, L2, link=%d, unit=1>" % ("double" if args.dtype == "float64" else "float", link)
        traffic = measured_traffic(("f64" if args.dtype == "float64" else "f32") + ("_link" if link else ""), nvox_loc)
        nzg = main_res["nz_global"]
        line = {
            "metric": METRIC if (head_scaling == "strong" or world == 1) else METRIC_WEAK, "value": main_res["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": head_scaling, "vs_baseline": None,
            "dtype": "f64" if args.dtype == "float64" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (nzg, n, n, ALPHA, L2, args.iters), "input": INPUT,
                       "parallelism": ("z-slab x%d (%d planes per GPU, %d z-chunks per launch), 3-plane halo exchange per iteration, %s" % (
                           world, nzg // world, main_res["chunks"],
                           "inside the iteration kernels over peer memory (NVLink), no host work per iteration"
                           if halo_mode[0] == "p2p" else "NCCL send/recv")) if world > 1 else "single GPU",
                       "l2_policy": "inputs larger than L2 (%.1f GiB of solver state per GPU)" % (main_res["plan_bytes"] / 2.0 ** 30)},
            "gpu_launches": main_res["launches"],
            "clocks": main_res["clocks"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["bytes"] if traffic else None, "traffic_source": traffic["source"] if traffic else None,
                         "peak_source": peak_src, "kernel": kernel_name,
                         "algorithmic_bytes_per_launch": 11 * esz * nvox_loc, "avg_launch_ms": main_res["iter_ms"],
                         "copy_gbs_this_gpu": copy_gbs, "frac_of_copy_this_gpu": achieved / copy_gbs,
                         "frac_of_nominal_8000_gbs": achieved / 8000.0},
        }
        if "e2e" in main_res:
            line["e2e"] = main_res["e2e"]
            line["checksum"] = main_res.get("checksum")
        if "parity" in main_res:
            line["parity"] = main_res["parity"]
        if weak_res is not None:
            key = "weak" if head_scaling == "strong" else "strong"
            sub = {"metric": METRIC_WEAK if key == "weak" else METRIC, "value": weak_res["value"], "unit": UNIT,
                   "ms_per_step": weak_res["ms_per_step"], "steps": weak_res["steps"],
                   "workload": WORKLOAD % (weak_res["nz_global"], n, n, ALPHA, L2, args.iters),
                   "roofline_frac": 11 * esz * weak_res["nvox_loc"] / (weak_res["iter_ms"] * 1e-3) / 1e9 / peak}
            if "e2e" in weak_res:
                sub["e2e"] = {k: weak_res["e2e"][k] for k in ("value", "unit", "ms_per_step", "steps", "h2d_bytes_per_step", "d2h_bytes_per_step")}
            line[key] = sub
        if admm_rec is not None:
            line["admm_slab"] = admm_rec
        if other is not None:
            oesz = 8 if esz == 4 else 4
            oach = 11 * oesz * other["nvox_loc"] / (other["iter_ms"] * 1e-3) / 1e9
            line["other_dtype"] = {"dtype": "f32" if oesz == 4 else "f64", "value": other["value"], "iter_ms": other["iter_ms"],
                                   "roofline_frac": oach / peak, "achieved_gbs": oach}
        if world == 1 and not args.no_extras:
            for key, fn in (("other_configs", other_configs), ("lsmr_roofline", lsmr_roofline_records)):
                try:
                    line[key] = fn(ctx, stream, peak)
                except Exception as e:      # never lose the headline line
                    line[key] = {"error": "%s: %s" % (type(e).__name__, e)}
        if not args.no_cpu_baseline and world == 1:
            try:
                rec = cpu_baseline_record(args.ref_size, 20.0, steps=1)
                line["cpu_baseline"] = {k: rec[k] for k in ("value", "unit", "cores", "kind", "host_cores", "sample", "versions")}
            except Exception as e:
                line["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
