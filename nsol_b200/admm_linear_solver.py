"""ADMM for TV-regularised linear least squares on the GPU.

API of ``nsol.admm_linear_solver.ADMMLinearSolver`` (nsol/admm_linear_solver.py:28-312):

    min_x 1/2 ||A x - b||^2 + alpha TV_iso(B x),   B = gradient

x-update: Tikhonov/LSMR solve on [A; sqrt(rho) B] with b_reg = v - w (cold start, clipped
to [0, inf)); v-update: isotropic soft threshold with alpha/rho; scaled dual w.  The whole
outer/inner loop runs on the device (``nsol_admm_run_host``) without host synchronisation.
"""
import numpy as np

from nsol_b200.linear_solver import LinearSolver, acquire_lsmr_plan


class ADMMLinearSolver(LinearSolver):

    def __init__(self, A, A_adj, b, B, B_adj, x0, dimension, b_reg=0, alpha=0.01, iter_max=10, minimizer="lsmr",
                 data_loss="linear", data_loss_scale=1, rho=0.5, iterations=10, x_scale=1, verbose=0, dtype=None):
        LinearSolver.__init__(self, A=A, A_adj=A_adj, b=b, x0=x0, alpha=alpha, iter_max=iter_max,
                              minimizer=minimizer, data_loss=data_loss, data_loss_scale=data_loss_scale,
                              x_scale=x_scale, verbose=verbose, dtype=dtype)
        self._B = B
        self._B_adj = B_adj
        self._b_reg = b_reg / self._x_scale
        self._dimension = dimension
        self._rho = float(rho)
        self._iterations = iterations
        self._dist = None

    def distribute(self, group=None):
        """Shard ONE tall volume over the ranks of an initialised ``torch.distributed`` group (additive API; the
        reference is single-process).  Every rank constructs the solver on ITS z-slab (contiguous planes along
        numpy axis 0, rank order = slab order; A / B built for the slab's shape, the same ``x_scale`` everywhere)
        and all ranks call ``run()`` together: blur halos travel on a ring (periodic boundary), gradient halos
        between neighbours, the LSMR norms are all-reduced (nsol_b200/distributed.py: SlabADMM).  ``get_x()``
        returns the rank's slab of the solution."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("ADMMLinearSolver.distribute: torch.distributed is not initialised")
        self.release()
        self._dist = {"group": group, "rank": dist.get_rank(group), "world": dist.get_world_size(group)}
        return self

    def release(self):
        slab = getattr(self, "_slab_admm", None)
        if slab is not None:
            slab.close()
            self._slab_admm = None
        LinearSolver.release(self)

    def set_rho(self, rho):
        self._rho = rho

    def get_rho(self):
        return self._rho

    def get_dimension(self):
        return self._dimension

    def set_iterations(self, iterations):
        self._iterations = iterations

    def get_iterations(self):
        return self._iterations

    def _get_cost_regularization_term(self, x):
        # isotropic total variation (nsol/prior_measures.py:27-37)
        parts = np.array_split(self._B(x), self._dimension)
        acc = parts[0] ** 2
        for p in parts[1:]:
            acc += p ** 2
        return np.sum(np.sqrt(acc))

    def _run(self):
        self._check_lsmr_only()
        info = self._probe_lsq(self._B, self._B_adj)
        if info["b_kind"] != "grad":
            raise TypeError("ADMMLinearSolver: B must be a LinearOperators gradient operator")
        if info["dim"] != self._dimension:
            raise ValueError("ADMMLinearSolver: dimension=%d but B is a %dD gradient" % (self._dimension, info["dim"]))
        n = self._x0.size
        iters = int(self._iterations)
        if self._dist is not None:
            return self._run_distributed(info, iters)
        plan = acquire_lsmr_plan(self, info, self._dtype)       # kept across runs (parameter studies)
        ctx = plan.ctx
        b = np.ascontiguousarray(self._b, dtype=np.float64)
        x0 = np.ascontiguousarray(self._x0, dtype=np.float64)   # v = B(x0) (:171); lsmr itself is cold-started
        # the solver's own b_reg (already divided by x_scale, :100): scalar or dim*N values
        b_reg = None
        if np.ndim(self._b_reg) != 0 or float(self._b_reg) != 0.0:
            b_reg = np.ascontiguousarray(np.broadcast_to(np.asarray(self._b_reg, dtype=np.float64).reshape(-1)
                                                         if np.ndim(self._b_reg) else self._b_reg, (info["dim"] * n,)),
                                         dtype=np.float64)
        ctx.check(ctx.lib.nsol_admm_set_b_reg_host(plan.handle, b_reg.ctypes.data if b_reg is not None else None, 1.0, None))
        x_out = ctx.result_empty(n, np.float64)
        its = np.empty((iters + 1, n), dtype=np.float64) if self._observer is not None else None
        ctx.check(ctx.lib.nsol_admm_run_host(
            plan.handle, float(self._alpha), float(self._rho), iters, int(self._iter_max), 1.0,
            float(self._x_scale), b.ctypes.data, x0.ctypes.data, x_out.ctypes.data,
            its.ctypes.data if its is not None else None, None))
        if its is not None:
            # nsol/admm_linear_solver.py:168-169, 186-187
            for i in range(iters + 1):
                self._observer.add_x(np.array(its[i]))
        self._set_result(x_out)

    def _run_distributed(self, info, iters):
        """z-slab sharded run (see ``distribute``)."""
        import torch
        from nsol_b200 import _lib
        from nsol_b200.distributed import SlabADMM
        if np.ndim(self._b_reg) != 0 or float(self._b_reg) != 0.0:
            raise ValueError("ADMMLinearSolver.distribute: b_reg != 0 is not available in the sharded run")
        if self._observer is not None:
            raise TypeError("ADMMLinearSolver.distribute: an Observer would have to gather every iterate; not supported")
        if info["dim"] < 2:
            raise ValueError("ADMMLinearSolver.distribute: z-slab sharding needs a 2-D or 3-D grid")
        key = (tuple(info["shape"]), tuple(info["spacing"]), info["a_kind"], _lib.dtype_code(self._dtype))
        slab = getattr(self, "_slab_admm", None)
        if slab is None or self._slab_key != key:
            self.release()
            device = torch.device("cuda", torch.cuda.current_device())
            slab = SlabADMM(_lib.context(), info, self._dtype, self._dist["rank"], self._dist["world"], device,
                            group=self._dist["group"])
            self._slab_admm, self._slab_key = slab, key
        x = slab.run(np.asarray(self._b, dtype=np.float64), np.asarray(self._x0, dtype=np.float64), float(self._alpha),
                     float(self._rho), iters, int(self._iter_max))
        self._set_result(x * self._x_scale)
