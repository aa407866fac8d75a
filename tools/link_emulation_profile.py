#!/usr/bin/env python
"""Two linked 64-plane z-slabs of 512 x 512 on ONE GPU, driven alternately on one stream (the emulation of the in-kernel halo
exchange the tests use) -- something ncu can capture: `ncu -k regex:pd_iter_bulk ... python tools/link_emulation_profile.py`.
With --unlinked the same two slabs iterate on their own (zero boundary), for comparison."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from nsol_b200 import _lib

unlinked = "--unlinked" in sys.argv
ctx = _lib.context()
lib = ctx.lib
shape = (64, 512, 512)
rng = np.random.RandomState(1)
alpha = np.array([0.05])
plans, blocks = [], []
for r in range(2):
    desc = _lib.PdDesc()
    desc.grid = _lib.make_grid(shape, None, _lib.F64, 1)
    desc.reg, desc.data, desc.alg = _lib.REG["TV"], _lib.DATA["L2"], _lib.ALG["ALG2"]
    desc.huber_gamma, desc.L2 = 0.05, 8.0
    desc.x_scale = desc.x0_scale = desc.b_scale = 255.0
    desc.alpha = alpha.ctypes.data_as(_lib.c_double_p)
    h = C.c_void_p()
    ctx.check(lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(h)))
    plans.append(h)
    if not unlinked:
        blk, nbytes = C.c_void_p(), C.c_size_t()
        ctx.check(lib.nsol_pd_plan_link_create(h, C.byref(blk), C.byref(nbytes)))
        blocks.append(blk)
if not unlinked:
    ctx.check(lib.nsol_pd_plan_link_connect(plans[0], None, blocks[1]))
    ctx.check(lib.nsol_pd_plan_link_connect(plans[1], blocks[0], None))
for h in plans:
    obs = rng.rand(int(np.prod(shape))) * 255
    ctx.check(lib.nsol_pd_plan_reset_host(h, obs.ctypes.data, None, None))
for h in plans:
    ctx.check(lib.nsol_pd_plan_iterate(h, 0, None))          # publish the start states
for _ in range(6):
    for h in plans:
        ctx.check(lib.nsol_pd_plan_iterate(h, 1, None))
out = np.empty(int(np.prod(shape)))
ctx.check(lib.nsol_pd_plan_get_x_host(plans[0], out.ctypes.data, None))
print("linked" if not unlinked else "unlinked", "checksum", float(out[::1001].sum()))
for h in plans:
    lib.nsol_pd_plan_destroy(h)
