#!/usr/bin/env python
"""Multi-GPU parity check of the PUBLIC API: PrimalDualSolver.distribute() and ADMMLinearSolver.distribute().

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/check_api_multi_gpu.py [--shape 96 64 128] [--iters 25]

Every rank constructs the reference-shaped solver on ITS z-slab of one volume, calls distribute() and run(); rank 0
gathers get_x() of all ranks and compares with the unsharded solver run on its own GPU: the primal-dual result must
be bit-identical (float64), the ADMM result within 1e-12 (the all-reduced norms are summed in a different order).
Prints one JSON line; exits non-zero on a mismatch.  With NSOL_PD_PIPE=1 (and NSOL_PD_PIPE_PLANES / NSOL_PD_PIPE_DEPTH) in the
environment the primal-dual solves take the pipelined host solve through the linked slabs even at small sizes.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=3, default=[96, 64, 128])
    ap.add_argument("--iters", type=int, default=25)
    ap.add_argument("--halo", default="auto")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import nsol_b200.admm_linear_solver as admm
    import nsol_b200.linear_operators as lo
    import nsol_b200.primal_dual_solver as pd
    from nsol_b200.distributed import slab_bounds
    from nsol_b200.proximal_operators import ProximalOperators as prox

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    shape = tuple(args.shape)
    rng = np.random.RandomState(21)
    vol = rng.rand(*shape) * 255.0            # identical on every rank
    xs = float(vol.max())
    z_lo, z_hi = slab_bounds(shape[0], rank, world)

    def pd_solver(obs):
        sh = obs.shape
        grad, grad_adj = lo.LinearOperators3D().get_gradient_operators()
        zsh = (3 * sh[0],) + sh[1:]
        b = obs.flatten()
        return pd.PrimalDualSolver(prox_f=lambda x, tau: prox.prox_ell2_denoising(x, tau, x0=b, x_scale=xs), prox_g_conj=prox.prox_tv_conj,
                                   B=lambda x: grad(x.reshape(*sh)).flatten(), B_conj=lambda x: grad_adj(x.reshape(*zsh)).flatten(),
                                   L2=8, x0=b, alpha=0.05, iterations=args.iters, x_scale=xs)

    def admm_solver(obs):
        sh = obs.shape
        ops = lo.LinearOperators3D()
        A, A_adj = ops.get_gaussian_blurring_operators(np.eye(3))
        grad, grad_adj = ops.get_gradient_operators()
        zsh = (3 * sh[0],) + sh[1:]
        w = lambda op, s: (lambda x: op(x.reshape(*s)).flatten())
        return admm.ADMMLinearSolver(A=w(A, sh), A_adj=w(A_adj, sh), b=obs.flatten(), B=w(grad, sh), B_adj=w(grad_adj, zsh), x0=obs.flatten(),
                                     dimension=3, alpha=0.01, rho=0.1, iterations=3, iter_max=6, x_scale=xs)

    out = {"world": world, "shape": shape}
    ok = True
    for name, make, tol in (("primal_dual", pd_solver, 0.0), ("admm", admm_solver, 1e-12)):
        s = make(np.ascontiguousarray(vol[z_lo:z_hi]))
        if name == "primal_dual":
            s.distribute(halo=args.halo)
        else:
            s.distribute()
        s.run()
        s.run()                                # a second run on the same sharded solver (generation counters, cached plan)
        mine = np.array(s.get_x())
        if name == "primal_dual":
            import ctypes as C
            from nsol_b200 import _lib
            g, d = C.c_int(), C.c_int()
            _lib.context().lib.nsol_pd_plan_solve_info(s._plan, C.byref(g), C.byref(d))
            out["pipelined_groups_depth_rank%d" % rank] = [g.value, d.value]      # (0, 0): plain upload / iterate / download
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
        s.release()
        if rank == 0:
            ref = make(vol)
            ref.run()
            xr = ref.get_x()
            got = np.concatenate(parts)
            err = float(np.max(np.abs(got - xr)) / np.max(np.abs(xr)))
            out[name] = {"rel_max_abs": err, "bit_identical": bool(np.array_equal(got, xr)), "tolerance": tol}
            ok = ok and err <= tol
            ref.release()
        dist.barrier()
    if rank == 0:
        out["ok"] = ok
        print(json.dumps(out), flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
