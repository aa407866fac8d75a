"""ctypes binding of libnsol_b200.so (include/nsol_b200.h).

There is no CPU fallback: importing this module is cheap, but the first call
that needs the library raises ``RuntimeError`` if the shared object has not
been built (``python -m nsol_b200.build``) or no CUDA device is present.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnsol_b200.so")

NSOL_OK, NSOL_EINVAL, NSOL_ECUDA, NSOL_ENOMEM, NSOL_ESTATE, NSOL_ENCCL = 0, -1, -2, -3, -4, -5
F64, F32 = 0, 1
REG = {"TV": 0, "HUBER": 1, "TK1": 2}
DATA = {"L1": 0, "L2": 1}
ALG = {"ALG2": 0, "ALG2_AHMOD": 1, "ALG3": 2}
B_GRAD, B_IDENTITY, B_NONE = 0, 1, 2
PROX = {"TV_CONJ": 0, "HUBER_CONJ": 1, "TK1_CONJ": 2, "ELL1": 3, "ELL2": 4}
A_BLUR, A_IDENTITY = 0, 1
SLAB_V, SLAB_U0, SLAB_UZ, SLAB_X = 0, 1, 2, 3
(PH_RHS, PH_SCAL_INIT_BETA, PH_ADJ_FIRST, PH_SCAL_INIT_ALPHA, PH_FWD, PH_SCAL_BETA, PH_ADJ, PH_SCAL_ALPHA, PH_UPDATE,
 PH_SCAL_TESTS, PH_CLIP, PH_ADMM_INIT, PH_ADMM_SHRINK) = range(13)

c_void_pp = C.POINTER(C.c_void_p)
c_double_p = C.POINTER(C.c_double)


class Grid(C.Structure):
    _fields_ = [("dim", C.c_int32), ("dtype", C.c_int32), ("shape", C.c_int64 * 3),
                ("spacing", C.c_double * 3), ("batch", C.c_int32), ("reserved", C.c_int32)]


class PdDesc(C.Structure):
    _fields_ = [("grid", Grid), ("reg", C.c_int32), ("data", C.c_int32), ("alg", C.c_int32),
                ("b_batched", C.c_int32), ("huber_gamma", C.c_double), ("L2", C.c_double),
                ("x_scale", C.c_double), ("x0_scale", C.c_double), ("b_scale", C.c_double),
                ("alpha", c_double_p)]


class LsqDesc(C.Structure):
    _fields_ = [("grid", Grid), ("a_op", C.c_int32), ("b_op", C.c_int32), ("taps", c_double_p * 3),
                ("radius", C.c_int32 * 3), ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol include/nsol_b200.h declares
SIGNATURES = {
    "nsol_version": (C.c_int, []),
    "nsol_create": (C.c_int, [C.c_int, c_void_pp]),
    "nsol_destroy": (None, [C.c_void_p]),
    "nsol_last_error": (C.c_char_p, [C.c_void_p]),
    "nsol_set_tuning": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "nsol_launch_count": (C.c_int64, [C.c_void_p]),
    "nsol_device_sm_count": (C.c_int, [C.c_void_p]),
    "nsol_debug_guard_check": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "nsol_device_alloc": (C.c_int, [C.c_void_p, C.c_size_t, c_void_pp]),
    "nsol_device_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsol_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, c_void_pp]),
    "nsol_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsol_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nsol_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nsol_memcpy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nsol_memset_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]),
    "nsol_stream_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsol_scale_convert": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_double, C.c_int, C.c_void_p]),
    "nsol_grad": (C.c_int, [C.c_void_p, C.POINTER(Grid), C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_grad_adj": (C.c_int, [C.c_void_p, C.POINTER(Grid), C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_diff": (C.c_int, [C.c_void_p, C.POINTER(Grid), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_blur_sep": (C.c_int, [C.c_void_p, C.POINTER(Grid), C.POINTER(c_double_p), C.POINTER(C.c_int32),
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_conv_wrap": (C.c_int, [C.c_void_p, C.POINTER(Grid), c_double_p, C.POINTER(C.c_int64),
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_prox_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                  C.c_double, C.c_void_p, C.c_void_p]),
    "nsol_similarity_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "nsol_prior_stats": (C.c_int, [C.c_void_p, C.POINTER(Grid), C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_create": (C.c_int, [C.c_void_p, C.POINTER(PdDesc), c_void_pp]),
    "nsol_pd_plan_destroy": (None, [C.c_void_p]),
    "nsol_pd_plan_update": (C.c_int, [C.c_void_p, C.POINTER(PdDesc)]),
    "nsol_pd_plan_bytes": (C.c_size_t, [C.c_void_p]),
    "nsol_pd_plan_reset_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_reset_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_iterate": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "nsol_pd_plan_iterations_done": (C.c_int, [C.c_void_p]),
    "nsol_pd_plan_x_dev": (C.c_int, [C.c_void_p, c_void_pp]),
    "nsol_pd_plan_get_x_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_get_x_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_solve_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_set_pipe_direction": (C.c_int, [C.c_void_p, C.c_int]),
    "nsol_pd_plan_solve_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nsol_pd_run_host": (C.c_int, [C.c_void_p, C.POINTER(PdDesc), C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_set_halo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_boundary_planes": (C.c_int, [C.c_void_p, c_void_pp, c_void_pp, c_void_pp]),
    "nsol_pd_plan_iterate_part": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "nsol_pd_plan_boundary_planes_next": (C.c_int, [C.c_void_p, c_void_pp, c_void_pp, c_void_pp]),
    "nsol_pd_plan_chunks": (C.c_int, [C.c_void_p]),
    "nsol_pd_plan_link_create": (C.c_int, [C.c_void_p, c_void_pp, C.POINTER(C.c_size_t)]),
    "nsol_pd_plan_link_ipc_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_link_open": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_link_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_pd_plan_link_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsol_lsmr_plan_create": (C.c_int, [C.c_void_p, C.POINTER(LsqDesc), c_void_pp]),
    "nsol_lsmr_plan_destroy": (None, [C.c_void_p]),
    "nsol_lsmr_plan_bytes": (C.c_size_t, [C.c_void_p]),
    "nsol_lsmr_solve_dev": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                      C.c_double, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "nsol_tikhonov_run_host": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "nsol_admm_run_host": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_admm_set_b_reg_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "nsol_pd_deconv_run_host": (C.c_int, [C.c_void_p, C.POINTER(PdDesc), C.c_int, C.c_int, C.c_double, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_admm_run_dev": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsol_lsmr_plan_slab": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "nsol_lsmr_slab_buffers": (C.c_int, [C.c_void_p, C.c_int, c_void_pp, c_void_pp, c_void_pp, c_void_pp, C.POINTER(C.c_int)]),
    "nsol_lsmr_slab_arrays": (C.c_int, [C.c_void_p, c_void_pp, c_void_pp, c_void_pp]),
    "nsol_lsmr_slab_phase": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p]),
    "nsol_lsmr_plan_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "nsol_admm_shrink": (C.c_int, [C.c_void_p, C.POINTER(Grid), C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def load():
    """dlopen the shared library and declare every prototype.  No GPU needed."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "libnsol_b200.so is not built (%s missing); run `python -m nsol_b200.build`. "
                    "nsol_b200 has no CPU fallback." % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, (restype, argtypes) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


def _raise(rc, message):
    if rc == NSOL_EINVAL:
        raise ValueError(message)
    if rc == NSOL_ENOMEM:
        raise MemoryError(message)
    raise RuntimeError(message)


class Context(object):
    """One nsol_ctx per (process, device)."""

    def __init__(self, device=-1):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.nsol_create(device, C.byref(h))
        if rc != NSOL_OK:
            _raise(NSOL_ECUDA if rc != NSOL_EINVAL else rc,
                   (self.lib.nsol_last_error(None) or b"nsol_create failed").decode())
        self.handle = h
        self._pool = {}          # nbytes -> [pinned pointers ready for reuse]
        self._pool_bytes = 0

    def check(self, rc):
        if rc != NSOL_OK:
            _raise(rc, (self.lib.nsol_last_error(self.handle) or b"").decode() or "nsol error %d" % rc)

    def set_tuning(self, key, value):
        self.check(self.lib.nsol_set_tuning(self.handle, key.encode(), int(value)))

    def launch_count(self):
        return int(self.lib.nsol_launch_count(self.handle))

    def sm_count(self):
        return int(self.lib.nsol_device_sm_count(self.handle))

    def guard_check(self):
        """(guard bytes overwritten so far, guarded arrays alive) -- see the "debug_guard" tuning knob."""
        bad, n = C.c_int64(), C.c_int()
        self.check(self.lib.nsol_debug_guard_check(self.handle, C.byref(bad), C.byref(n)))
        return int(bad.value), int(n.value)

    # ---- memory
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.nsol_device_alloc(self.handle, nbytes, C.byref(p)))
        return DeviceBuffer(self, p, nbytes)

    def pinned_empty(self, shape, dtype=np.float64):
        """numpy array backed by page-locked host memory (freed with the array)."""
        shape = tuple(int(s) for s in np.atleast_1d(shape))
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.check(self.lib.nsol_host_alloc(self.handle, max(nbytes, 1), C.byref(p)))
        owner = _PinnedOwner(self, p)
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        arr = _PinnedArray(arr, owner)
        return arr

    # Results of large solves come back into page-locked buffers: a pageable np.empty() of 1 GiB costs
    # ~1 s of page faults + staged copies per get_x(), a pinned one ~20 ms.  cudaHostAlloc itself is slow,
    # so buffers whose array has been garbage-collected are kept for the next result of the same size.
    POOL_MIN_BYTES = 1 << 20
    POOL_MAX_BYTES = 8 << 30

    def result_empty(self, shape, dtype=np.float64):
        """Fresh host array for a solver result: page-locked (pooled) when large, plain numpy otherwise."""
        shape = tuple(int(s) for s in np.atleast_1d(shape))
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if nbytes < self.POOL_MIN_BYTES:
            return np.empty(shape, dtype=dtype)
        with _lock:
            free = self._pool.get(nbytes)
            ptr = free.pop() if free else None
            if ptr is not None:
                self._pool_bytes -= nbytes
        if ptr is None:
            p = C.c_void_p()
            self.check(self.lib.nsol_host_alloc(self.handle, nbytes, C.byref(p)))
            ptr = p.value
        owner = _PooledOwner(self, ptr, nbytes)
        buf = (C.c_char * nbytes).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return _PinnedArray(arr, owner)

    def _pool_return(self, ptr, nbytes):
        with _lock:
            if self.handle and self._pool_bytes + nbytes <= self.POOL_MAX_BYTES:
                self._pool.setdefault(nbytes, []).append(ptr)
                self._pool_bytes += nbytes
                return
        self.lib.nsol_host_free(self.handle, C.c_void_p(ptr))

    def pool_trim(self):
        """Release every cached page-locked buffer."""
        with _lock:
            ptrs = [p for lst in self._pool.values() for p in lst]
            self._pool.clear()
            self._pool_bytes = 0
        for p in ptrs:
            self.lib.nsol_host_free(self.handle, C.c_void_p(p))

    def sync(self, stream=None):
        self.check(self.lib.nsol_stream_sync(self.handle, stream))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.pool_trim()
                self.lib.nsol_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _PinnedOwner(object):
    def __init__(self, ctx, ptr):
        self.ctx, self.ptr = ctx, ptr

    def __del__(self):
        try:
            self.ctx.lib.nsol_host_free(self.ctx.handle, self.ptr)
        except Exception:
            pass


class _PooledOwner(object):
    """Hands its page-locked buffer back to the context's pool when the last array view dies."""

    def __init__(self, ctx, ptr, nbytes):
        self.ctx, self.ptr, self.nbytes = ctx, ptr, nbytes

    def __del__(self):
        try:
            self.ctx._pool_return(self.ptr, self.nbytes)
        except Exception:
            pass


class _PinnedArray(np.ndarray):
    """ndarray view that keeps its pinned allocation alive."""

    def __new__(cls, arr, owner):
        obj = arr.view(cls)
        obj._nsol_owner = owner
        return obj

    def __array_finalize__(self, obj):
        # only views of the page-locked buffer keep it alive; arrays numpy allocates for results derived from
        # this one (ufunc outputs, copies: base is None) must not pin it
        self._nsol_owner = getattr(obj, "_nsol_owner", None) if self.base is not None else None

    def __array_wrap__(self, out_arr, context=None, return_scalar=False):
        # results of arithmetic on a solver result are plain ndarrays, as with the reference
        out = np.asarray(out_arr).view(np.ndarray) if isinstance(out_arr, np.ndarray) else out_arr
        return out[()] if return_scalar else out


class DeviceBuffer(object):
    def __init__(self, ctx, ptr, nbytes):
        self.ctx, self.ptr, self.nbytes = ctx, ptr, nbytes

    def upload(self, arr, stream=None):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.ctx.check(self.ctx.lib.nsol_memcpy_h2d(self.ctx.handle, self.ptr, arr.ctypes.data, arr.nbytes, stream))
        self.ctx.sync(stream)   # arr may be a temporary
        return self

    def download(self, shape, dtype, stream=None):
        out = self.ctx.result_empty(shape, dtype)
        assert out.nbytes <= self.nbytes
        self.ctx.check(self.ctx.lib.nsol_memcpy_d2h(self.ctx.handle, out.ctypes.data, self.ptr, out.nbytes, stream))
        self.ctx.sync(stream)
        return out

    def free(self):
        if self.ptr is not None:
            self.ctx.lib.nsol_device_free(self.ctx.handle, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


_contexts = {}


def context(device=None):
    """Process-wide context for ``device`` (default: the current CUDA device)."""
    key = -1 if device is None else int(device)
    with _lock:
        ctx = _contexts.get(key)
    if ctx is None:
        ctx = Context(key)
        with _lock:
            _contexts[key] = ctx
    return ctx


def current_context():
    """The already-created context of the current device, or None (never creates one)."""
    with _lock:
        return _contexts.get(-1) or (next(iter(_contexts.values())) if _contexts else None)


def make_grid(shape, spacing=None, dtype=F64, batch=1):
    shape = tuple(int(s) for s in shape)
    dim = len(shape)
    if dim < 1 or dim > 3:
        raise ValueError("only 1-, 2- and 3-dimensional grids are supported (got %d)" % dim)
    g = Grid()
    g.dim = dim
    g.dtype = dtype
    g.batch = batch
    sp = np.ones(dim) if spacing is None else np.atleast_1d(np.asarray(spacing, dtype=np.float64))
    if sp.size != dim:
        raise ValueError("dimension of spacing and space must be the same")
    for a in range(3):
        g.shape[a] = shape[a] if a < dim else 1
        g.spacing[a] = float(sp[a]) if a < dim else 1.0
    return g


def dtype_code(dtype):
    if dtype in (F64, "float64", "f64", np.float64, None):
        return F64
    if dtype in (F32, "float32", "f32", np.float32):
        return F32
    raise ValueError("dtype must be 'float64' or 'float32' (got %r)" % (dtype,))


def np_dtype(code):
    return np.float32 if code == F32 else np.float64
