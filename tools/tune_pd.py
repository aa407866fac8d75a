#!/usr/bin/env python
"""Sweep the tuning knobs of the fused primal-dual kernel on one GPU (device-resident input).
    python tools/tune_pd.py [--size 512] [--iters 20] [--dtype float64] [--ty 4,8,16] [--zc 16,32,64]
Prints one line per configuration: ms per iteration and fraction of the measured HBM peak."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nsol_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--nz", type=int, default=0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--ty", default="0")
    ap.add_argument("--zc", default="0")
    ap.add_argument("--variant", default="0")
    ap.add_argument("--reg", default="TV")
    ap.add_argument("--data", default="L2")
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--warm", type=int, default=3, help="untimed iterations before the timed ones (use ~300 for the power-capped sustained regime)")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    ctx = _lib.context(0)
    n = args.size
    nz = args.nz or n
    shape = (nz, n, n) if args.dim == 3 else (n, n)
    nvox = int(np.prod(shape))
    dcode = _lib.dtype_code(args.dtype)
    esz = 4 if dcode == _lib.F32 else 8
    tdt = torch.float32 if esz == 4 else torch.float64
    obs = torch.rand(nvox, dtype=tdt, device="cuda") * 255
    desc = _lib.PdDesc()
    desc.grid = _lib.make_grid(shape, None, dcode, args.batch)
    desc.reg, desc.data, desc.alg = _lib.REG[args.reg], _lib.DATA[args.data], _lib.ALG["ALG2"]
    desc.huber_gamma, desc.L2 = 0.05, 8.0
    desc.x_scale = desc.x0_scale = desc.b_scale = 255.0
    alphas = np.linspace(0.01, 0.05, args.batch)
    desc.alpha = alphas.ctypes.data_as(_lib.c_double_p)
    plan = C.c_void_p()
    ctx.check(ctx.lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(plan)))
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    words = 5 + 2 * args.dim
    for variant in [int(v) for v in args.variant.split(",")]:
        for ty in [int(v) for v in args.ty.split(",")]:
            for zc in [int(v) for v in args.zc.split(",")]:
                ctx.set_tuning("pd_ty", ty)
                ctx.set_tuning("pd_zc", zc)
                ctx.set_tuning("pd_variant", variant)
                ctx.check(ctx.lib.nsol_pd_plan_reset_dev(plan, C.c_void_p(obs.data_ptr()), None, stream))
                ctx.check(ctx.lib.nsol_pd_plan_iterate(plan, args.warm, stream))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.check(ctx.lib.nsol_pd_plan_iterate(plan, args.iters, stream))
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.iters
                gbs = words * esz * nvox * args.batch / (ms * 1e-3) / 1e9
                print("variant=%d ty=%2d zc=%3d  %.4f ms/iter  %.3e vox-it/s  %.0f GB/s  %.3f of peak"
                      % (variant, ty, zc, ms, nvox * args.batch / (ms * 1e-3), gbs, gbs / peak), flush=True)
    ctx.lib.nsol_pd_plan_destroy(plan)


if __name__ == "__main__":
    main()
