"""Host <-> device plumbing for the operator closures: run one C-ABI operator
on numpy input.  Arithmetic always happens on the GPU (no CPU fallback)."""
import ctypes as C

import numpy as np

from nsol_b200 import _lib


def run_array_op(call, x, out_shape, n_tmp=0):
    """Upload float64 ``x``, run ``call(ctx, x_dev, out_dev, tmp_dev)``, download ``out_shape``."""
    ctx = _lib.context()
    x = np.ascontiguousarray(x, dtype=np.float64)
    n_out = int(np.prod(out_shape))
    d_in = ctx.device_alloc(max(x.nbytes, 8))
    d_out = ctx.device_alloc(max(n_out * 8, 8))
    d_tmp = ctx.device_alloc(max(n_tmp * 8, 8)) if n_tmp else None
    try:
        d_in.upload(x)
        call(ctx, d_in.ptr, d_out.ptr, d_tmp.ptr if d_tmp else None)
        return d_out.download(out_shape, np.float64)
    finally:
        d_in.free()
        d_out.free()
        if d_tmp:
            d_tmp.free()


def as_double_p(arr):
    return arr.ctypes.data_as(_lib.c_double_p)
