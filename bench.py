#!/usr/bin/env python
"""Benchmark of the NSoL proximal-solver hot path on B200 (BASELINE.json metric):
voxel-iterations/s of the fused 3-D TV-L2 primal-dual (Chambolle-Pock) iteration at 512^3.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype float64|float32]
                    [--size 512] [--iters 100] [--scaling weak|strong] [--halo auto|p2p|nccl]
                    [--other-configs] [--secondary-dtype] [--impl reference]

One "step" = one full solve: reset (x = xbar = b/x_scale, p = 0) + ``iters`` fused
primal-dual iterations over the volume (BASELINE config 4: alpha=0.05, L2=8, ALG2).
  value      device-timed (CUDA events on the launching stream, max over ranks) with the
             observation resident in HBM;
  e2e        the same solve through the public API (PrimalDualSolver.run() + get_x() at
             N=1; the z-slab driver at N>1) with pinned HOST buffers, copies in the timed region;
  roofline   algorithmic bytes (5+2d words per voxel-iteration) / measured
             average launch duration of the iteration kernel vs MEASURED_PEAKS.json;
  cpu_baseline  the oracle port of the reference loop (same scipy.ndimage / numpy call
             structure as nsol/primal_dual_solver.py:232-261) on a bounded sample.
Multi-GPU (torchrun, one rank per GPU): z-slab decomposition with a 3-plane halo
exchange per iteration, by default inside the iteration kernels over CUDA-IPC mapped peer
memory (--halo p2p; NCCL send/recv with --halo nccl; nsol_b200/distributed.py).
``--impl reference`` times the CPU port alone on rank 0 (the reference is pure Python and
cannot travel to the GPU box).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALPHA, L2 = 0.05, 8.0
METRIC = "voxel-iterations/s, 3D TV-L2 primal-dual (Chambolle-Pock ALG2), 512^3"


def load_phantom():
    z = np.load(os.path.join(ROOT, "tests", "golden", "inputs.npz"))
    return z["shepp_logan_64"].astype(np.float64)


def synth_volume(shape, seed=1):
    """BASELINE config 4 input: the reference's 64^3 Shepp-Logan fixture repeated along each
    axis to ``shape`` plus Noise(seed).add_gaussian_noise(0.05) (nsol/noise.py:51-55)."""
    ph = load_phantom()
    reps = [int(np.ceil(s / 64.0)) for s in shape]
    vol = ph
    for ax, r in enumerate(reps):
        vol = np.repeat(vol, r, axis=ax) if r > 1 else vol
    # np.repeat stretches each voxel; crop to the requested shape
    vol = vol[:shape[0], :shape[1], :shape[2]]
    rng = np.random.RandomState(seed)
    return vol + 0.05 * 255.0 * rng.standard_normal(size=vol.shape)


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


def measured_traffic(dtype_key, nvox):
    """DRAM bytes per launch of the iteration kernel from the committed ncu capture, scaled to the
    number of voxels of this launch (the capture is at 512^3)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as fh:
            t = json.load(fh)[dtype_key]
        return t["bytes_per_launch"] * (nvox / float(t["voxels"]))
    except Exception:
        return None


def device_copy_bandwidth(nbytes=1 << 31):
    """Device-to-device copy bandwidth (read + write bytes) of THIS GPU in this run, measured the way
    MEASURED_PEAKS.json was (torch b.copy_(a), best of 10): separates GPU-to-GPU variation of the HBM
    peak from the kernel's own efficiency."""
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


WORKLOAD = "C4: 3D TV-L2 primal-dual denoising %dx%dx%d, alpha=%g, L2=%g, ALG2, %d iterations per step"


def cpu_reference_throughput(sample_n, iters):
    """The reference's CPU loop (oracle port, reference call structure) on a sample_n^3 sub-volume."""
    from oracle import nsol_oracle as orc
    vol = synth_volume((sample_n, sample_n, sample_n))
    b = vol.reshape(-1)
    t0 = time.perf_counter()
    orc.primal_dual_denoise_ndimage(b, vol.shape, reg="TV", data="L2", alpha=ALPHA, L2=L2, iterations=iters,
                                    x_scale=float(vol.max()))
    dt = time.perf_counter() - t0
    return vol.size * iters / dt, dt


def run_reference_arm(args, rank):
    """CPU arm: the oracle port of the reference's own loop, timed on the host cores of rank 0."""
    if rank != 0:
        return
    n = args.ref_size
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_throughput(min(n, 64), 1)
    times = []
    for _ in range(args.steps):
        thr, dt = cpu_reference_throughput(n, args.ref_iters)
        times.append(dt)
    dt = float(np.mean(times))
    value = n ** 3 * args.ref_iters / dt
    sample = ("%d^3 sub-volume of the config-4 input, %d iterations per step; numpy/scipy.ndimage port of "
              "nsol/primal_dual_solver.py:232-261 (the reference loop is single-threaded)" % (n, args.ref_iters))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxel-iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % (args.size * (args.gpus if args.scaling == "weak" else 1), args.size, args.size,
                                           ALPHA, L2, args.iters),
                   "input": "64^3 Shepp-Logan fixture repeated to size + Gaussian noise 0.05 (seeded)",
                   "cpu_sample": "%d^3 sub-volume, %d iterations per step" % (n, args.ref_iters)},
        "cpu_baseline": {"value": value, "unit": "voxel-iterations/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": "voxel-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def other_configs(ctx, stream, peak):
    """Short device-timed runs of the other BASELINE configurations (1, 2, 3, 5) on one GPU:
    latency-bound 2-D single images, the ADMM/LSMR deconvolution and the batched alpha sweep.
    Inputs are the reference's test images (tests/golden/inputs.npz) with seeded synthetic noise."""
    import torch
    from nsol_b200 import _lib
    import nsol_b200.linear_operators as lo
    import nsol_b200.admm_linear_solver as admm
    z = np.load(os.path.join(ROOT, "tests", "golden", "inputs.npz"))
    rng = np.random.RandomState(1)
    lib = ctx.lib
    out = {}

    def time_events(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

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
                "hbm_frac_of_measured": words * esz * vox_it / (ms * 1e-3) / 1e9 / peak, "dtype": dtype,
                "note": "includes the H2D copy of the observation and the reset"}

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
        "ms_per_solve": dt * 1e3, "pixel_lsmr_iters_per_s": img.size * 50 * 10 / dt,
        "note": "ADMMLinearSolver.run()+get_x() wall clock (host buffers; outer iterations replayed as a CUDA graph)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--iters", type=int, default=100, help="primal-dual iterations per step (solve)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--ref-size", type=int, default=256)
    ap.add_argument("--ref-iters", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--overlap", action="store_true",
                    help="multi-GPU: split every iteration (boundary chunks, exchange || interior chunks); measured no faster "
                         "than the plain exchange on 8 B200 (profiles/r1_scaling.md), so off by default")
    ap.add_argument("--halo", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU halo exchange: p2p = inside the iteration kernels over peer memory (NVLink), "
                         "nccl = grouped send/recv before every launch; auto = p2p if CUDA IPC works")
    ap.add_argument("--secondary-dtype", action="store_true", help="also time the other dtype (reported under 'other_dtype')")
    ap.add_argument("--other-configs", action="store_true", help="also time BASELINE configs 1, 2, 3, 5 (reported under 'other_configs')")
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
    # global volume and the local z-slab
    nz_global = n * world if args.scaling == "weak" else n
    z_lo, z_hi = slab_bounds(nz_global, rank, world)
    nz_loc = z_hi - z_lo
    shape = (nz_loc, n, n)
    nvox_loc = nz_loc * n * n
    nvox_global = nz_global * n * n

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

    # ---- synthetic observation, pinned on the host, then resident in HBM -------------
    obs_host = ctx.pinned_empty((nvox_loc,), np.float64)
    vol = synth_volume(shape, seed=1 + rank)
    x_scale = max_over_ranks(float(vol.max()))   # x_scale = max(observed) (nsol/application/run_denoising.py:97)
    obs_host[:] = vol.reshape(-1)
    del vol
    out_host = ctx.pinned_empty((nvox_loc,), np.float64)

    halo_mode = ["single"]

    def measure(dtype_name, want_e2e):
        dcode = _lib.dtype_code(dtype_name)
        esz = 4 if dcode == _lib.F32 else 8
        np_dt = _lib.np_dtype(dcode)
        desc = _lib.PdDesc()
        desc.grid = _lib.make_grid(shape, None, dcode, 1)
        desc.reg, desc.data, desc.alg = _lib.REG["TV"], _lib.DATA["L2"], _lib.ALG["ALG2"]
        desc.b_batched = 0
        desc.huber_gamma, desc.L2 = 0.05, L2
        desc.x_scale = desc.x0_scale = desc.b_scale = x_scale
        alpha_arr = np.array([ALPHA])
        desc.alpha = alpha_arr.ctypes.data_as(_lib.c_double_p)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        slab = SlabPrimalDual(ctx, desc, n * n, np_dt, rank, world, device, halo=args.halo)
        halo_mode[0] = slab.mode
        lib, plan = ctx.lib, slab.plan
        # observation resident in HBM in the plan's dtype
        stage = ctx.device_alloc(nvox_loc * 8)
        obs_dev = ctx.device_alloc(nvox_loc * esz)
        ctx.check(lib.nsol_memcpy_h2d(ctx.handle, stage.ptr, obs_host.ctypes.data, nvox_loc * 8, stream))
        ctx.check(lib.nsol_scale_convert(ctx.handle, nvox_loc, _lib.F64, stage.ptr, dcode, obs_dev.ptr, 1.0, 0, stream))
        torch.cuda.synchronize()
        stage.free()

        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

        def solve_resident():
            slab.reset_dev(obs_dev.ptr, None, stream)
            ev[1].record()
            slab.iterate(args.iters, stream, overlap=args.overlap)
            ev[2].record()

        for _ in range(args.warmup):
            solve_resident()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = ctx.launch_count()
        kernel_ms = 0.0
        barrier()
        ev[0].record()
        pairs = []
        for _ in range(args.steps):
            e1, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            slab.reset_dev(obs_dev.ptr, None, stream)
            e1.record()
            slab.iterate(args.iters, stream, overlap=args.overlap)
            e2.record()
            pairs.append((e1, e2))
        ev[3].record()
        barrier()
        slab.check(stream)      # a timed-out in-kernel halo wait invalidates the run
        total_ms = max_over_ranks(ev[0].elapsed_time(ev[3]))
        kernel_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in pairs))
        launches = ctx.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        res = {
            "value": nvox_global * args.iters * args.steps / (total_ms * 1e-3),
            "ms_per_step": total_ms / args.steps,
            "iter_ms": kernel_ms / (args.steps * args.iters),
            "launches": launches,
            "clocks": clocks,
            "bytes_per_voxel_iter": 11 * esz,
            "plan_bytes": int(lib.nsol_pd_plan_bytes(plan)),
        }
        # ---- end to end through the public API, host buffers ---------------------------
        if want_e2e:
            slab.close()
            obs_dev.free()
            if world == 1:
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
                def e2e_step():
                    solver.run()
                    return solver.get_x()
            else:
                slab2 = SlabPrimalDual(ctx, desc, n * n, np_dt, rank, world, device, halo=args.halo)
                def e2e_step():
                    slab2.reset_host(obs_host.ctypes.data, None, stream)
                    slab2.iterate(args.iters, stream, overlap=args.overlap)
                    ctx.check(lib.nsol_pd_plan_get_x_host(slab2.plan, out_host.ctypes.data, stream))
                    return out_host
            # warm-up: two results alive at once, so the page-locked result pool holds a spare buffer
            # and no timed step pays for a cudaHostAlloc
            x = e2e_step()
            x_spare = e2e_step()
            del x_spare
            barrier()
            link = pcie_bandwidth(ctx, stream) if rank == 0 else None
            barrier()
            t0 = time.perf_counter()
            n_e2e = max(1, min(args.steps, 3))
            for _ in range(n_e2e):
                x = e2e_step()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            res["e2e"] = {"value": nvox_global * args.iters * n_e2e / dt, "unit": "voxel-iterations/s",
                          "h2d_bytes_per_step": 2 * nvox_loc * 8 if world == 1 else nvox_loc * 8,
                          "d2h_bytes_per_step": nvox_loc * 8, "ms_per_step": dt / n_e2e * 1e3, "steps": n_e2e,
                          "api": "PrimalDualSolver.run()+get_x()" if world == 1 else
                                 "SlabPrimalDual: nsol_pd_plan_reset_host + iterate + nsol_pd_plan_get_x_host",
                          "host_link": link}
            res["checksum"] = float(np.sum(x[:: max(1, x.size // 4096)]))
            if world > 1:
                slab2.close()
        else:
            slab.close()
            obs_dev.free()
        return res

    main_res = measure(args.dtype, not args.no_e2e)
    other = None
    if args.secondary_dtype:
        od = "float32" if args.dtype == "float64" else "float64"
        other = measure(od, False)

    peak, peak_src = measured_peak()
    copy_gbs = device_copy_bandwidth() if rank == 0 else None
    if rank == 0:
        esz = 4 if args.dtype == "float32" else 8
        achieved = 11 * esz * nvox_loc / (main_res["iter_ms"] * 1e-3) / 1e9
        # default variants (csrc/pd_kernels.cu): TMA bulk-async kernel for float64 3-D, register-pipelined LDG for float32
        variant = int(os.environ.get("NSOL_PD_VARIANT", "0")) or (2 if args.dtype == "float64" else 1)
        kernel_name = ("pd_iter_bulk_kernel" if variant == 2 else "pd_iter_kernel") + \
            "<%s, TV, L2, link=%d, unit=1>" % ("double" if args.dtype == "float64" else "float", 1 if halo_mode[0] == "p2p" else 0)
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": "voxel-iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if args.dtype == "float64" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (nz_global, n, n, ALPHA, L2, args.iters),
                       "input": "64^3 Shepp-Logan fixture repeated to size + Gaussian noise 0.05 (seeded)",
                       "parallelism": ("z-slab x%d, 3-plane halo exchange per iteration, %s" % (
                           world, "inside the iteration kernels over peer memory (NVLink), no host work per iteration"
                           if halo_mode[0] == "p2p" else
                           "NCCL send/recv" + (" overlapped with the interior chunks" if args.overlap else ""))) if world > 1 else "single GPU",
                       "l2_policy": "inputs larger than L2 (%.1f GiB of solver state per GPU)" % (main_res["plan_bytes"] / 2.0 ** 30)},
            "gpu_launches": main_res["launches"],
            "clocks": main_res["clocks"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic("f64" if args.dtype == "float64" else "f32", nvox_loc), "peak_source": peak_src, "kernel": kernel_name,
                         "algorithmic_bytes_per_launch": 11 * esz * nvox_loc, "avg_launch_ms": main_res["iter_ms"],
                         "copy_gbs_this_gpu": copy_gbs, "frac_of_copy_this_gpu": achieved / copy_gbs,
                         "frac_of_nominal_8000_gbs": achieved / 8000.0},
        }
        if "e2e" in main_res:
            line["e2e"] = main_res["e2e"]
        if other is not None:
            oesz = 8 if esz == 4 else 4
            oach = 11 * oesz * nvox_loc / (other["iter_ms"] * 1e-3) / 1e9
            line["other_dtype"] = {"dtype": "f32" if oesz == 4 else "f64", "value": other["value"], "iter_ms": other["iter_ms"],
                                   "roofline_frac": oach / peak, "achieved_gbs": oach}
        if args.other_configs and world == 1:
            try:
                line["other_configs"] = other_configs(ctx, C.c_void_p(torch.cuda.current_stream().cuda_stream), peak)
            except Exception as e:      # never lose the headline line
                line["other_configs"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if not args.no_cpu_baseline and world == 1:
            thr, dt = cpu_reference_throughput(args.ref_size, args.ref_iters)
            line["cpu_baseline"] = {
                "value": thr, "unit": "voxel-iterations/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
                "sample": "%d^3 sub-volume of the same input, %d iterations (%.1f s); numpy/scipy.ndimage port with the "
                          "reference's call structure, single-threaded like the reference" % (args.ref_size, args.ref_iters, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
