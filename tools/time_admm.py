#!/usr/bin/env python
"""Time the device-resident ADMM TV-L2 deconvolution (nsol_admm_run_dev) for a given image size.
    python tools/time_admm.py [--size 512] [--dim 2] [--iterations 50] [--iter-max 10] [--dtype float64]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nsol_b200 import _lib  # noqa: E402
from nsol_b200 import kernels as K  # noqa: E402
from nsol_b200.linear_solver import LsmrPlan  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--iter-max", type=int, default=10)
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    shape = (args.size,) * args.dim
    n = int(np.prod(shape))
    mask = getattr(K, "Kernels%dD" % args.dim)().get_gaussian(np.eye(args.dim) if args.dim > 1 else 1.0)

    class Op(object):
        taps = K.separable_taps(mask)
    info = dict(a_kind="conv", a_op=Op, b_kind="grad", shape=shape, spacing=(1.0,) * args.dim, dim=args.dim)
    plan = LsmrPlan(info, args.dtype)
    ctx = plan.ctx
    tdt = torch.float32 if args.dtype == "float32" else torch.float64
    b = torch.rand(n, dtype=tdt, device="cuda")
    x = torch.empty_like(b)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        ctx.check(ctx.lib.nsol_admm_run_dev(plan.handle, 0.01, 0.1, args.iterations, args.iter_max, C.c_void_p(b.data_ptr()),
                                            C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), stream))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    inner = args.iterations * args.iter_max
    words = 19 if args.dim == 2 else (22 if args.dim == 3 else 16)
    esz = 4 if args.dtype == "float32" else 8
    print("path=%s blocks=%s shape=%s %s: %.3f ms per solve, %.2f us per inner iteration, %.3e px-LSMR-it/s, %.0f GB/s algorithmic (%d words)"
          % (os.environ.get("NSOL_LSMR_PATH", "auto"), os.environ.get("NSOL_LSMR_BLOCKS", "auto"), shape, args.dtype, ms,
             ms * 1e3 / inner, n * inner / (ms * 1e-3), words * esz * n * inner / (ms * 1e-3) / 1e9, words), flush=True)
    plan.close()


if __name__ == "__main__":
    main()
