#!/usr/bin/env python
"""Multi-GPU parity + timing of the z-slab ADMM TV-L2 deconvolution (one rank per GPU, NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
        tools/check_admm_multi_gpu.py [--shape 96 64 80] [--iterations 4] [--iter-max 6] [--dtype float64] [--no-check]

Every rank runs its slab through nsol_b200.distributed.SlabADMM (ring exchange of the blur halos,
neighbour exchange of the gradient halos, all-reduce of the LSMR norms); rank 0 gathers the slabs and
compares with the UNSHARDED solve of the same volume on its own GPU (ADMMLinearSolver of the reference
API).  Tolerance 1e-9 relative in float64 (the norm reductions are summed in a different order).
Prints one JSON line; exits non-zero on a mismatch.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs="+", default=[96, 64, 80])
    ap.add_argument("--iterations", type=int, default=4)
    ap.add_argument("--iter-max", type=int, default=6)
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--alpha", type=float, default=0.02)
    ap.add_argument("--rho", type=float, default=0.2)
    ap.add_argument("--no-check", action="store_true", help="timing only (volume too large for one GPU / too slow unsharded)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from nsol_b200 import _lib
    import nsol_b200.linear_operators as lo
    import nsol_b200.admm_linear_solver as admm
    from nsol_b200.distributed import SlabADMM, slab_bounds
    from nsol_b200.linear_solver import probe_least_squares

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    ctx = _lib.context(local_rank)

    shape = tuple(args.shape)
    dim = len(shape)
    n = int(np.prod(shape))
    z_lo, z_hi = slab_bounds(shape[0], rank, world)
    # every rank generates only its own planes (seeded per plane, so the volume does not depend on the world size)
    mine = np.stack([np.random.RandomState(1000 + z).rand(*shape[1:]) * 200.0 + 20.0 for z in range(z_lo, z_hi)])
    xs = 220.0
    ops = getattr(lo, "LinearOperators%dD" % dim)()
    A, A_adj = ops.get_gaussian_blurring_operators(np.eye(dim))
    grad, grad_adj = ops.get_gradient_operators()
    zshape = (dim * shape[0],) + shape[1:]
    fA = lambda x: A(x.reshape(*shape)).flatten()
    fA_adj = lambda x: A_adj(x.reshape(*shape)).flatten()
    fD = lambda x: grad(x.reshape(*shape)).flatten()
    fD_adj = lambda x: grad_adj(x.reshape(*zshape)).flatten()
    info = probe_least_squares(fA, fA_adj, fD, fD_adj, n)
    local = dict(info, shape=(z_hi - z_lo,) + shape[1:])
    solver = SlabADMM(ctx, local, args.dtype, rank, world, device)
    part = mine.reshape(-1) / xs

    out = solver.run(part, part, args.alpha, args.rho, 1, 2)          # warm-up (NCCL channels, allocations)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    out = solver.run(part, part, args.alpha, args.rho, args.iterations, args.iter_max)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    inner = args.iterations * args.iter_max
    result = {"world": world, "shape": list(shape), "dtype": args.dtype, "iterations": args.iterations, "iter_max": args.iter_max,
              "seconds": dt, "voxel_lsmr_iters_per_s": n * inner / dt,
              "note": "wall clock of SlabADMM.run incl. H2D of the slab and D2H of the result"}
    ok = True
    if not args.no_check:
        t = torch.from_numpy(out * xs).to(device)
        sizes = [(slab_bounds(shape[0], r, world)[1] - slab_bounds(shape[0], r, world)[0]) * (n // shape[0]) for r in range(world)]
        if rank == 0:
            parts = [torch.empty(sz, dtype=torch.float64, device=device) for sz in sizes]
            parts[0] = t
            for r in range(1, world):
                dist.recv(parts[r], src=r)
            sharded = torch.cat(parts).cpu().numpy()
            full = np.stack([np.random.RandomState(1000 + z).rand(*shape[1:]) * 200.0 + 20.0 for z in range(shape[0])]).reshape(-1)
            ref_solver = admm.ADMMLinearSolver(A=fA, A_adj=fA_adj, b=full, B=fD, B_adj=fD_adj, x0=full, dimension=dim, alpha=args.alpha,
                                               rho=args.rho, iterations=args.iterations, iter_max=args.iter_max, x_scale=xs, dtype=args.dtype)
            ref_solver.run()
            ref = ref_solver.get_x()
            diff = float(np.max(np.abs(sharded - ref)) / np.max(np.abs(ref)))
            tol = 1e-9 if args.dtype == "float64" else 2e-3
            ok = diff < tol
            result.update(rel_max=diff, tol=tol)
        else:
            dist.send(t, dst=0)
    solver.close()
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.broadcast(flag, src=0)
    if rank == 0:
        result["ok"] = bool(ok)
        print(json.dumps(result))
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
