#!/usr/bin/env python
"""BASELINE config 5 on N GPUs: 64-point alpha sweep x {TV, Huber, TK1} primal-dual denoising of a
1024^2 image (200 iterations each), the 192 independent runs dealt round-robin to the ranks with no
data-path communication (PrimalDualSolverParameterStudy under torch.distributed, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
        tools/run_sweep_multi_gpu.py [--size 1024] [--points 64] [--iterations 200] [--dtype float64]

Rank 0 prints one JSON line: wall time of the three studies (including the npz/txt writers), runs/s and
voxel-iterations/s, and -- parity -- the max relative deviation of 3 sampled reconstructions from the
same points solved one by one on rank 0 (must be 0 in float64: batching and fan-out do not change bits;
the study stores float16 reconstructions like the reference, so the sample is compared after the same cast).
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--points", type=int, default=64)
    ap.add_argument("--iterations", type=int, default=200)
    ap.add_argument("--dtype", default="float64")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import nsol_b200.linear_operators as lo
    import nsol_b200.primal_dual_solver as pd
    import nsol_b200.primal_dual_solver_parameter_study as pdps
    from nsol_b200.proximal_operators import ProximalOperators as prox
    from nsol_b200.observer import Observer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)     # host-side gather of results only

    z = np.load(os.path.join(ROOT, "tests", "golden", "inputs.npz"))
    img = z["man_1024"].astype(np.float64)[:args.size, :args.size]
    rng = np.random.RandomState(1)
    obs = img + 0.05 * img.max() * rng.standard_normal(img.shape)
    shape = obs.shape
    xs = float(obs.max())
    grad, grad_adj = lo.LinearOperators2D().get_gradient_operators()
    zshape = (2 * shape[0], shape[1])
    D = lambda x: grad(x.reshape(*shape)).flatten()
    D_adj = lambda x: grad_adj(x.reshape(*zshape)).flatten()
    b = obs.flatten()
    prox_f = lambda x, tau: prox.prox_ell2_denoising(x, tau, x0=b, x_scale=xs)
    regs = {"TV": prox.prox_tv_conj, "HUBER": prox.prox_huber_conj, "TK1": (lambda q, s: q / (1 + s))}
    alphas = np.linspace(0.001, 0.05, args.points)
    base = tempfile.mkdtemp(prefix="nsol_sweep_") if rank == 0 else None
    if world > 1:
        box = [base]
        dist.broadcast_object_list(box, src=0)
        base = box[0]

    def study(reg):
        solver = pd.PrimalDualSolver(prox_f=prox_f, prox_g_conj=regs[reg], B=D, B_conj=D_adj, L2=8, x0=b, alpha=alphas[0],
                                     iterations=args.iterations, x_scale=xs, dtype=args.dtype)
        st = pdps.PrimalDualSolverParameterStudy(solver, Observer(), dir_output=os.path.join(base, reg), name=reg,
                                                 parameters={"alpha": alphas}, reconstruction_info={"shape": shape})
        st.run()
        return solver

    # where the wall clock goes: time inside run_sweep (H2D, iterations, D2H) and inside the npz writer
    import nsol_b200.solver_parameter_study as sps
    spent = {"run_sweep": 0.0, "savez": 0.0}
    orig_sweep, orig_savez = pd.PrimalDualSolver.run_sweep, sps.npz_members

    def timed_sweep(self, alphas_):
        t = time.perf_counter()
        out = orig_sweep(self, alphas_)
        spent["run_sweep"] += time.perf_counter() - t
        return out

    def timed_savez(*a, **k):
        t = time.perf_counter()
        out = orig_savez(*a, **k)
        spent["savez"] += time.perf_counter() - t
        return out

    pd.PrimalDualSolver.run_sweep = timed_sweep
    sps.npz_members = timed_savez
    study("TV")      # warm-up: plan allocation, page-locked pools
    spent["run_sweep"] = spent["savez"] = 0.0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for reg in ("TV", "HUBER", "TK1"):
        study(reg)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        runs = 3 * args.points
        line = {"workload": "C5: %d-point alpha sweep x {TV, Huber, TK1} PD on %dx%d, %d iterations" % (args.points, shape[0], shape[1], args.iterations),
                "n_gpus": world, "dtype": args.dtype, "seconds": dt, "runs": runs, "runs_per_s": runs / dt,
                "voxel_iters_per_s": runs * obs.size * args.iterations / dt,
                "seconds_in_run_sweep_rank0": spent["run_sweep"], "seconds_in_npz_compression_rank0": spent["savez"],
                "note": "wall clock of three PrimalDualSolverParameterStudy.run() calls incl. result gather and npz/txt writers"}
        # parity sample: three points solved one at a time
        worst = 0.0
        for reg, idx in (("TV", 0), ("HUBER", args.points // 2), ("TK1", args.points - 1)):
            rec = np.load(os.path.join(base, reg, reg + "_reconstructions.npz"))   # float16, keyed by the run index
            s = pd.PrimalDualSolver(prox_f=prox_f, prox_g_conj=regs[reg], B=D, B_conj=D_adj, L2=8, x0=b, alpha=float(alphas[idx]),
                                    iterations=args.iterations, x_scale=xs, dtype=args.dtype)
            s.run()
            one = s.get_x().astype(np.float16).astype(np.float64)
            got = np.asarray(rec[str(idx)], dtype=np.float64).reshape(-1)
            worst = max(worst, float(np.max(np.abs(got - one)) / np.max(np.abs(one))))
        line["parity_sample_rel_max"] = worst
        line["ok"] = worst == 0.0 if args.dtype == "float64" else worst < 1e-3
        print(json.dumps(line))
        shutil.rmtree(base, ignore_errors=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
