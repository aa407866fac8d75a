#!/usr/bin/env python
"""Debug aid: the linked pipelined host solve on one GPU (one context / stream / thread per slab); dumps the link headers."""
import ctypes as C
import sys
import threading

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from nsol_b200 import _lib
from nsol_b200.distributed import slab_bounds

nslabs, nz, planes, depth = [int(v) for v in sys.argv[1:5]]
iters_list = [int(v) for v in sys.argv[5:]] or [13, 6]
shape = (nz, 10, 68)
rng = np.random.RandomState(nz)
obs = rng.rand(*shape) * 255
xs = float(obs.max())
alpha = np.array([0.05])
ctxs = [_lib.Context(-1) for _ in range(nslabs)]
streams = [torch.cuda.Stream() for _ in range(nslabs)]
lib = ctxs[0].lib
plans, spans, blocks = [], [], []
for r, ctx in enumerate(ctxs):
    for key, val in (("pd_zc", 2), ("pd_pipe", 1), ("pd_pipe_planes", planes), ("pd_pipe_depth", depth), ("link_timeout_ms", 3000)):
        ctx.set_tuning(key, val)
    z_lo, z_hi = slab_bounds(shape[0], r, nslabs)
    spans.append((z_lo, z_hi))
    desc = _lib.PdDesc()
    desc.grid = _lib.make_grid((z_hi - z_lo,) + shape[1:], None, _lib.F64, 1)
    desc.reg, desc.data, desc.alg = _lib.REG["TV"], _lib.DATA["L2"], _lib.ALG["ALG2"]
    desc.huber_gamma, desc.L2 = 0.05, 8.0
    desc.x_scale = desc.x0_scale = desc.b_scale = xs
    desc.alpha = alpha.ctypes.data_as(_lib.c_double_p)
    h = C.c_void_p()
    ctx.check(lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(h)))
    plans.append(h)
    blk, nbytes = C.c_void_p(), C.c_size_t()
    ctx.check(lib.nsol_pd_plan_link_create(h, C.byref(blk), C.byref(nbytes)))
    blocks.append(blk)
for r, (ctx, h) in enumerate(zip(ctxs, plans)):
    ctx.check(lib.nsol_pd_plan_link_connect(h, blocks[r - 1] if r > 0 else None, blocks[r + 1] if r < nslabs - 1 else None))
    ctx.check(lib.nsol_pd_plan_set_pipe_direction(h, -1 if r % 2 else 1))
slabs = [np.ascontiguousarray(obs[z_lo:z_hi]).reshape(-1) for z_lo, z_hi in spans]


def dump(tag):
    torch.cuda.synchronize()
    for r in range(nslabs):
        hdr = np.zeros(64, dtype=np.uint32)
        ctxs[r].check(lib.nsol_memcpy_d2h(ctxs[r].handle, hdr.ctypes.data_as(C.c_void_p), blocks[r], 256, None))
        ctxs[r].check(lib.nsol_stream_sync(ctxs[r].handle, None))
        print(tag, "rank", r, "flag_below", hdr[0], "flag_above", hdr[16], "count_below", hdr[32], "count_above", hdr[48], "error", hdr[56], flush=True)


for iters in iters_list:
    outs = [np.empty(s.size) for s in slabs]
    errs = [None] * nslabs

    def work(r):
        try:
            ctxs[r].check(lib.nsol_pd_plan_solve_host(plans[r], slabs[r].ctypes.data, None, iters, outs[r].ctypes.data,
                                                      C.c_void_p(streams[r].cuda_stream)))
        except Exception as e:
            errs[r] = e
    threads = [threading.Thread(target=work, args=(r,)) for r in range(nslabs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    print("iters", iters, "errors", errs, flush=True)
    dump("after %d" % iters)
