#!/usr/bin/env python
"""Multi-GPU parity check of the z-slab primal-dual iteration (one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_slab_multi_gpu.py [--halo p2p|nccl|auto] [--shape 96 64 128] [--iters 25] [--dtype float64]

Every rank runs its slab (halo exchange inside the kernels over peer memory, or NCCL send/recv);
rank 0 gathers the slabs and compares them with the UNSHARDED run of the same volume on its own
GPU: bit for bit in float64, 1e-5 relative in float32.  Two solves back to back (the second with
a different iteration count) exercise the generation counters of the in-kernel exchange across a
reset.  Prints one JSON line and exits non-zero on a mismatch.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--halo", default="p2p", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--shape", type=int, nargs=3, default=[96, 64, 128])
    ap.add_argument("--iters", type=int, default=25)
    ap.add_argument("--dtype", default="float64")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from nsol_b200 import _lib
    from nsol_b200.distributed import SlabPrimalDual, slab_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    ctx = _lib.context(local_rank)
    lib = ctx.lib
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    shape = tuple(args.shape)
    rng = np.random.RandomState(11)
    obs = rng.rand(*shape) * 255.0            # identical on every rank
    xs = float(obs.max())
    dcode = _lib.dtype_code(args.dtype)
    np_dt = _lib.np_dtype(dcode)
    alpha = np.array([0.05])

    def make_desc(local_shape):
        desc = _lib.PdDesc()
        desc.grid = _lib.make_grid(local_shape, None, dcode, 1)
        desc.reg, desc.data, desc.alg = _lib.REG["TV"], _lib.DATA["L2"], _lib.ALG["ALG2"]
        desc.huber_gamma, desc.L2 = 0.05, 8.0
        desc.x_scale = desc.x0_scale = desc.b_scale = xs
        desc.alpha = alpha.ctypes.data_as(_lib.c_double_p)
        return desc

    z_lo, z_hi = slab_bounds(shape[0], rank, world)
    local_shape = (z_hi - z_lo,) + shape[1:]
    plane = shape[1] * shape[2]
    slab = SlabPrimalDual(ctx, make_desc(local_shape), plane, np_dt, rank, world, device, halo=args.halo)
    mine = np.ascontiguousarray(obs[z_lo:z_hi]).reshape(-1)
    result = {"halo": slab.mode, "world": world, "shape": list(shape), "dtype": args.dtype, "solves": []}
    ok = True
    for iters in (args.iters, max(1, args.iters // 3)):
        slab.reset_host(mine.ctypes.data, None, stream)
        slab.iterate(iters, stream)
        out = np.empty(mine.size)
        ctx.check(lib.nsol_pd_plan_get_x_host(slab.plan, out.ctypes.data, stream))
        slab.check(stream)
        # gather the slabs on rank 0
        sizes = [(slab_bounds(shape[0], r, world)[1] - slab_bounds(shape[0], r, world)[0]) * plane for r in range(world)]
        t = torch.from_numpy(out).to(device)
        if rank == 0:
            parts = [torch.empty(sz, dtype=torch.float64, device=device) for sz in sizes]
            parts[0] = t
            for r in range(1, world):
                dist.recv(parts[r], src=r)
            sharded = torch.cat(parts).cpu().numpy()
            # unsharded run on this GPU
            h = C.c_void_p()
            desc = make_desc(shape)
            ctx.check(lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(h)))
            flat = np.ascontiguousarray(obs.reshape(-1))
            ctx.check(lib.nsol_pd_plan_reset_host(h, flat.ctypes.data, None, stream))
            ctx.check(lib.nsol_pd_plan_iterate(h, iters, stream))
            ref = np.empty(flat.size)
            ctx.check(lib.nsol_pd_plan_get_x_host(h, ref.ctypes.data, stream))
            lib.nsol_pd_plan_destroy(h)
            diff = float(np.max(np.abs(sharded - ref)) / np.max(np.abs(ref)))
            exact = bool(np.array_equal(sharded, ref))
            good = exact if args.dtype == "float64" else diff < 1e-5
            ok = ok and good
            result["solves"].append({"iterations": iters, "bit_exact": exact, "rel_max": diff, "ok": good})
        else:
            dist.send(t, dst=0)
    slab.close()
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.broadcast(flag, src=0)
    if rank == 0:
        result["ok"] = ok
        print(json.dumps(result))
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
