"""First-order primal-dual (Chambolle-Pock) solver of the reference API, on the GPU.

Same constructor, setters/getters and ``run()/get_x()`` contract as
``nsol.primal_dual_solver.PrimalDualSolver`` (nsol/primal_dual_solver.py:26-403)
for  min_x [f(x) + alpha g(Bx)].  ``_run`` does not call the Python callables
per iteration: it probes them once (nsol_b200/_trace.py), recognises

    B, B_conj      = LinearOperators*.get_gradient_operators()      (possibly lambda-wrapped)
    prox_g_conj    = ProximalOperators.prox_tv_conj | prox_huber_conj | lambda q, s: q/(1+s)
    prox_f         = prox_ell1_denoising | prox_ell2_denoising  (x0=b, x_scale)   [denoising]
                   | prox_linear_least_squares (A = blur | identity)              [deconvolution]

and runs the whole loop nsol/primal_dual_solver.py:232-261 on the device through
the C ABI (``nsol_pd_*``): one fused kernel launch per iteration.  Callables that
are not built from these pieces raise ``TypeError`` -- there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from nsol_b200 import _lib
from nsol_b200 import _trace
from nsol_b200.solver import Solver, _memory_key

_ALG_TYPES = ("ALG2", "ALG2_AHMOD", "ALG3")


class PrimalDualSolver(Solver):

    def __init__(self, prox_f, prox_g_conj, B, B_conj, L2, x0, alpha=0.01, iterations=10, x_scale=1.,
                 verbose=0, alg_type="ALG2", dtype=None):
        Solver.__init__(self, x0=x0, verbose=verbose, x_scale=x_scale)
        self._prox_f = prox_f
        self._prox_g_conj = prox_g_conj
        self._B = B
        self._B_conj = B_conj
        self._L2 = float(L2)
        self._alpha = float(alpha)
        self._iterations = iterations
        self._alg_type = alg_type
        self._dtype = dtype          # additive option: "float64" (default) | "float32"
        self._config = None          # cached result of probing the callables
        self._dist = None            # z-slab sharding over a torch.distributed group (see distribute())

    # -- setters / getters (nsol/primal_dual_solver.py:120-197) ------------------
    def set_alpha(self, alpha):
        self._alpha = alpha

    def get_alpha(self):
        return self._alpha

    def set_L2(self, L2):
        self._L2 = L2

    def get_L2(self):
        return self._L2

    def set_alg_type(self, alg_type):
        self._alg_type = alg_type

    def get_alg_type(self):
        return self._alg_type

    def set_iterations(self, iterations):
        self._iterations = iterations

    def get_iterations(self):
        return self._iterations

    def set_dtype(self, dtype):
        self._dtype = dtype

    def get_dtype(self):
        return "float32" if _lib.dtype_code(self._dtype) == _lib.F32 else "float64"

    def distribute(self, group=None, halo="auto"):
        """Shard ONE tall volume over the ranks of an initialised ``torch.distributed`` process group (additive
        API; the reference is single-process).  Every rank constructs the solver on ITS z-slab -- contiguous
        planes along numpy axis 0, rank order = slab order: ``x0`` / the prox's observation are the slab, ``B`` /
        ``B_conj`` are gradient operators of the slab's shape, ``x_scale`` is the same on every rank -- and all
        ranks call ``run()`` together.  The halo planes travel inside the iteration kernels over NVLink peer
        memory (``halo="p2p"``) or by NCCL send/recv (``"nccl"``); ``get_x()`` returns the rank's slab of the
        solution, bit-identical to the corresponding planes of the unsharded solve."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("PrimalDualSolver.distribute: torch.distributed is not initialised")
        self.release()
        self._dist = {"group": group, "halo": halo, "rank": dist.get_rank(group), "world": dist.get_world_size(group)}
        return self

    def print_statistics(self, fmt="%.3e"):
        pass

    # -- probing -----------------------------------------------------------------
    def _probe(self):
        """Map the four callables to a kernel configuration (cached per solver)."""
        if self._config is not None:
            return self._config
        n = self._x0.size
        cfg = {}
        out = _trace.probe(self._B, n)
        if out.expr[0] != "grad" or out.expr[1] != ("arg",):
            raise TypeError("PrimalDualSolver: B must be a LinearOperators gradient operator; supported: "
                            + _trace.SUPPORTED)
        _, _, dim, spacing, shape = out.expr
        if int(np.prod(shape)) != n:
            raise ValueError("PrimalDualSolver: B reshapes x0 (%d values) to %s" % (n, (shape,)))
        cfg.update(dim=dim, spacing=spacing, shape=shape)
        adj = _trace.probe(self._B_conj, dim * n)
        if adj.expr[0] != "grad_adj" or adj.expr[1] != ("arg",) or adj.expr[2:5] != (dim, spacing, shape):
            raise TypeError("PrimalDualSolver: B_conj must be the adjoint gradient of B's grid; supported: "
                            + _trace.SUPPORTED)

        s1, s2 = 0.37, 1.83
        g1 = _trace.probe(self._prox_g_conj, dim * n, s1).expr
        g2 = _trace.probe(self._prox_g_conj, dim * n, s2).expr
        if g1[0] == "prox_tv_conj" and g1[1] == ("arg",):
            cfg.update(reg="TV", huber_gamma=0.05)
        elif g1[0] == "prox_huber_conj" and g1[1] == ("arg",) and g1[2] == s1 and g2[2] == s2:
            cfg.update(reg="HUBER", huber_gamma=g1[3])
        elif (g1[0] == "scale" and g1[1] == ("arg",) and g1[3] and g2[3]
              and g1[2] == 1 + s1 and g2[2] == 1 + s2):
            cfg.update(reg="TK1", huber_gamma=0.05)
        else:
            raise TypeError("PrimalDualSolver: unsupported prox_g_conj; supported: " + _trace.SUPPORTED)

        t1 = 0.4321
        f1 = _trace.probe(self._prox_f, n, t1).expr
        if f1[0] in ("prox_ell1", "prox_ell2") and f1[1] == ("arg",) and f1[2] == t1:
            b = np.asarray(f1[3], dtype=np.float64).reshape(-1)
            if b.size != n:
                raise ValueError("PrimalDualSolver: prox_f observation has %d values, x0 has %d" % (b.size, n))
            cfg.update(kind="denoise", data="L1" if f1[0] == "prox_ell1" else "L2", b=b, b_scale=f1[4],
                       b_src=_memory_key(f1[3]))
        elif f1[0] == "prox_lls" and f1[1] == ("arg",) and f1[2] == t1:
            cfg.update(kind="deconv", lls=f1)
        else:
            raise TypeError("PrimalDualSolver: unsupported prox_f; supported: " + _trace.SUPPORTED)
        self._config = cfg
        return cfg

    # -- execution ---------------------------------------------------------------
    def _make_desc(self, cfg, alphas):
        dtype = _lib.dtype_code(self._dtype)
        if self._alg_type not in _ALG_TYPES:
            raise KeyError(self._alg_type)
        desc = _lib.PdDesc()
        desc.grid = _lib.make_grid(cfg["shape"], cfg["spacing"], dtype, batch=len(alphas))
        desc.reg = _lib.REG[cfg["reg"]]
        desc.data = _lib.DATA[cfg["data"]]
        desc.alg = _lib.ALG[self._alg_type]
        desc.b_batched = 0
        desc.huber_gamma = cfg["huber_gamma"]
        desc.L2 = float(self._L2)
        desc.x_scale = float(self._x_scale)
        desc.x0_scale = 1.0            # self._x0 is already x0 / x_scale (nsol/solver.py:37)
        desc.b_scale = float(cfg["b_scale"])
        arr = np.ascontiguousarray(alphas, dtype=np.float64)
        desc.alpha = arr.ctypes.data_as(_lib.c_double_p)
        desc._keep = arr
        return desc

    def _x0_is_observation(self, cfg, b):
        """True when the caller's x0 was the very array the prox holds as its observation and the two scales agree
        (the reference's denoising wiring: ``b = x0 = observed.flatten()``, nsol/application/run_denoising.py:95-97).
        Then x0 / x_scale is evaluated on the device from the one uploaded copy of b (the same IEEE division the
        host did) and the second host->device copy -- 1 GiB at 512^3 -- is skipped.  A strided sample guards
        against the array having been modified since construction."""
        if cfg.get("b_src") is None or cfg["b_src"] != getattr(self, "_x0_src", None):
            return False
        if float(cfg["b_scale"]) != float(self._x_scale) or b.size != self._x0.size or b.size == 0:
            return False
        idx = np.linspace(0, b.size - 1, num=min(b.size, 64)).astype(np.int64)
        return bool(np.array_equal(np.asarray(self._x0)[idx], b[idx] / float(self._x_scale)))

    # -- device plan (kept between runs of the same solver: no reallocation in a sweep) ---------
    def _acquire_plan(self, ctx, cfg, desc):
        key = (cfg["shape"], cfg["spacing"], int(desc.grid.dtype), int(desc.grid.batch), None if self._dist is None else
               (self._dist["rank"], self._dist["world"], self._dist["halo"]))
        plan = getattr(self, "_plan", None)
        if plan is not None and self._plan_key == key:
            ctx.check(ctx.lib.nsol_pd_plan_update(plan, C.byref(desc)))
            return plan
        self.release()
        if self._dist is not None:
            # every rank creates its slab plan and exchanges link handles with its neighbours (collective)
            import torch
            from nsol_b200.distributed import SlabPrimalDual
            shape = cfg["shape"]
            if len(shape) < 2:
                raise ValueError("PrimalDualSolver.distribute: z-slab sharding needs a 2-D or 3-D grid")
            device = torch.device("cuda", torch.cuda.current_device())
            self._slab = SlabPrimalDual(ctx, desc, int(np.prod(shape[1:])), _lib.np_dtype(int(desc.grid.dtype)),
                                        self._dist["rank"], self._dist["world"], device, halo=self._dist["halo"],
                                        group=self._dist["group"])
            h = self._slab.plan
        else:
            h = C.c_void_p()
            ctx.check(ctx.lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(h)))
        self._plan, self._plan_key, self._plan_ctx = h, key, ctx
        return h

    def release(self):
        """Free the device memory held by this solver (plan + device-resident result; the LSMR plan of the
        deconvolution wiring).  With ``distribute()`` this is collective: every rank must call it."""
        from nsol_b200.linear_solver import release_lsmr_plan
        release_lsmr_plan(self)
        plan = getattr(self, "_plan", None)
        if plan is not None:
            if self._fetch_result is not None:
                self._x_unscaled = self._fetch_result()
                self._fetch_result = None
            slab = getattr(self, "_slab", None)
            if slab is not None:
                slab.close()
                self._slab = None
            else:
                self._plan_ctx.lib.nsol_pd_plan_destroy(plan)
            self._plan = None

    def __del__(self):
        try:
            from nsol_b200.linear_solver import release_lsmr_plan
            release_lsmr_plan(self)
            plan = getattr(self, "_plan", None)
            if plan is not None and getattr(self, "_slab", None) is None:
                self._plan_ctx.lib.nsol_pd_plan_destroy(plan)
                self._plan = None
        except Exception:
            pass

    def _run(self):
        cfg = self._probe()
        if cfg["kind"] == "deconv":
            if self._dist is not None:
                raise TypeError("PrimalDualSolver.distribute: only the denoising prox maps are sharded")
            from nsol_b200._pd_deconv import run_pd_deconvolution
            return run_pd_deconvolution(self, cfg)
        ctx = _lib.context()
        lib = ctx.lib
        n = self._x0.size
        desc = self._make_desc(cfg, [float(self._alpha)])
        b = np.ascontiguousarray(cfg["b"], dtype=np.float64)
        same = self._x0_is_observation(cfg, b)
        if same:
            desc.x0_scale = desc.b_scale        # x = xbar = b / x_scale, evaluated on the device from the copy of b
            x0 = None
        else:
            x0 = np.ascontiguousarray(self._x0, dtype=np.float64)
        iters = int(self._iterations)
        if iters < 0:
            raise ValueError("iterations must be >= 0")
        plan = self._acquire_plan(ctx, cfg, desc)
        slab = getattr(self, "_slab", None)

        def fetch():
            out = ctx.result_empty(n, np.float64)
            ctx.check(lib.nsol_pd_plan_get_x_host(plan, out.ctypes.data, None))
            return out

        if self._observer is None:
            # No Observer: the whole solve is ONE library call that also brings the result back -- for a large volume the upload,
            # the iterations and the download overlap (nsol_pd_plan_solve_host).  The first get_x() hands that array out, later
            # calls download a fresh copy like the reference returns a fresh array (nsol/solver.py:117-118).
            first = [ctx.result_empty(n, np.float64)]
            if slab is not None:
                slab.solve_host(b.ctypes.data, x0.ctypes.data if x0 is not None else None, iters, first[0].ctypes.data, None)
            else:
                ctx.check(lib.nsol_pd_plan_solve_host(plan, b.ctypes.data, x0.ctypes.data if x0 is not None else None, iters,
                                                      first[0].ctypes.data, None))

            def fetch_first():
                if first:
                    return first.pop()
                return fetch()
            self._set_device_result(fetch_first)
            return
        ctx.check(lib.nsol_pd_plan_reset_host(plan, b.ctypes.data, x0.ctypes.data if x0 is not None else None, None))
        if slab is not None:
            slab._halo_fresh = False

        def iterate(k):
            if slab is not None:
                slab.iterate(k, None)
            else:
                ctx.check(lib.nsol_pd_plan_iterate(plan, k, None))

        reqs = None
        if self._observer is not None and not getattr(self._observer, "get_store_iterates", lambda: True)():
            reqs = self._observer.device_measure_requests(n)
        if self._observer is None:
            iterate(iters)
        elif reqs is not None:
            if slab is not None:
                raise TypeError("PrimalDualSolver.distribute: device-side measures see the local slab only; use a storing Observer")
            # measures as device reductions on the resident iterate
            from nsol_b200.similarity_measures import device_stats, from_stats
            refs = {}
            for name, r in reqs.items():
                key = id(r.x_ref)
                if key not in refs:
                    xr = np.ascontiguousarray(np.asarray(r.x_ref, dtype=np.float64).reshape(-1))
                    if xr.size != n:
                        raise ValueError("measure '%s': reference has %d values, x has %d" % (name, xr.size, n))
                    refs[key] = ctx.device_alloc(xr.nbytes).upload(xr)
            results = {name: np.zeros(iters + 1) for name in reqs}
            xptr = C.c_void_p()
            dcode = int(desc.grid.dtype)
            try:
                for i in range(iters + 1):
                    if i > 0:
                        iterate(1)
                    ctx.check(lib.nsol_pd_plan_x_dev(plan, C.byref(xptr)))
                    cache = {}
                    for name, r in reqs.items():
                        key = id(r.x_ref)
                        if key not in cache:
                            cache[key] = device_stats(ctx, dcode, n, xptr, float(self._x_scale), refs[key])
                        results[name][i] = from_stats(r.kind, cache[key], n)
            finally:
                for buf in refs.values():
                    buf.free()
            self._observer.set_device_results(results)
            self._observer.add_x(fetch())
        else:
            # nsol/primal_dual_solver.py:218-219, 260-261: the observer sees x0 and every iterate
            self._observer.add_x(fetch())
            for _ in range(iters):
                iterate(1)
                self._observer.add_x(fetch())
        if slab is not None:
            slab.check(None)    # synchronises; a timed-out halo wait invalidates the result
        else:
            ctx.sync()          # run() returns when the solve is finished (computational time, errors)
        self._set_device_result(fetch)

    def run_sweep(self, alphas):
        """Batched parameter sweep: one fused launch per iteration advances every alpha.
        Returns an array (len(alphas), N) of get_x() results.  (Additive API used by the
        parameter-study driver; the reference runs the points one after another,
        nsol/solver_parameter_study.py:170-221.)"""
        cfg = self._probe()
        if cfg["kind"] != "denoise":
            raise TypeError("run_sweep is available for the denoising prox maps only")
        if self._x0.ndim != 1:
            raise ValueError("Initial value x0 must be a 1D array")
        ctx = _lib.context()
        n = self._x0.size
        alphas = [float(a) for a in alphas]
        desc = self._make_desc(cfg, alphas)
        x0 = np.ascontiguousarray(self._x0, dtype=np.float64)
        b = np.ascontiguousarray(cfg["b"], dtype=np.float64)
        x_out = ctx.result_empty((len(alphas), n), np.float64)
        ctx.check(ctx.lib.nsol_pd_run_host(ctx.handle, C.byref(desc), int(self._iterations), b.ctypes.data,
                                           x0.ctypes.data, x_out.ctypes.data, None, None))
        return x_out
