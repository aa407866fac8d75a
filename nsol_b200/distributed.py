"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink /
NVSwitch on the GPUs, gloo in the CPU tests).

The reference is single-process, so everything here is new.  Two ways
the path shards:

* independent units -- parameter-sweep points / image batches: ``partition_round_robin``;
  no data-path communication at all;
* one large volume -- z-slab decomposition of the primal-dual iteration: rank r owns the
  contiguous planes [z_lo, z_hi) of every array.  The fused iteration kernel needs, per
  iteration, from the upper neighbour the first xbar plane (forward difference at the
  slab top) and from the lower neighbour the last xbar and p_z planes (to recompute
  p'_z[z_lo-1] for the adjoint): 3 planes per interior boundary, exchanged with
  send/recv between neighbours.  The global ends keep the reference's zero (Dirichlet)
  boundary: no wrap-around.
"""
import numpy as np


def slab_bounds(nz, rank, world):
    """Contiguous balanced split of nz planes: returns (z_lo, z_hi) of ``rank``."""
    base, rem = divmod(int(nz), int(world))
    z_lo = rank * base + min(rank, rem)
    return z_lo, z_lo + base + (1 if rank < rem else 0)


def partition_round_robin(n_items, rank, world):
    """Indices of the independent work items (sweep points, images) owned by ``rank``."""
    return list(range(int(rank), int(n_items), int(world)))


class HaloExchanger(object):
    """Neighbour exchange of the three boundary planes of a z-slab.

    ``exchange(xbar_first, xbar_last, pz_last)`` sends this rank's boundary planes and
    fills ``xbar_above`` (from rank+1), ``xbar_below`` and ``pz_below`` (from rank-1).
    Tensors are torch tensors on the communicator's device.  Ranks at the global ends
    simply skip the missing neighbour (their halos stay unused: zero boundary)."""

    def __init__(self, rank, world, plane_numel, dtype, device, group=None):
        import torch
        self.torch = torch
        self.rank, self.world, self.group = rank, world, group
        self.has_below = rank > 0
        self.has_above = rank < world - 1
        mk = lambda: torch.zeros(plane_numel, dtype=dtype, device=device)
        self.xbar_above = mk() if self.has_above else None
        self.xbar_below = mk() if self.has_below else None
        self.pz_below = mk() if self.has_below else None

    def start_exchange(self, xbar_first, xbar_last, pz_last):
        """Enqueue the sends/receives (non-blocking for the host); returns the requests."""
        dist = self.torch.distributed
        ops = []
        if self.has_above:
            ops.append(dist.P2POp(dist.isend, xbar_last, self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.isend, pz_last, self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.xbar_above, self.rank + 1, self.group))
        if self.has_below:
            ops.append(dist.P2POp(dist.isend, xbar_first, self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.xbar_below, self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, self.pz_below, self.rank - 1, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish_exchange(reqs):
        for req in reqs:
            req.wait()      # NCCL: makes the current stream wait; gloo: blocks the host

    def exchange(self, xbar_first, xbar_last, pz_last):
        self.finish_exchange(self.start_exchange(xbar_first, xbar_last, pz_last))


class _DevicePtr(object):
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, numel, np_dtype):
        self.__cuda_array_interface__ = {
            "shape": (int(numel),), "typestr": np.dtype(np_dtype).str, "data": (int(ptr), False),
            "version": 2, "strides": None}


def tensor_from_ptr(ptr, numel, np_dtype, device):
    import torch
    return torch.as_tensor(_DevicePtr(ptr, numel, np_dtype), device=device)


class SlabPrimalDual(object):
    """z-slab sharded fused primal-dual iteration on one GPU per rank.

    Wraps an ``nsol_pd_plan`` built for the local slab.  Two ways to move the three halo planes:

    * ``halo="p2p"`` (default where CUDA IPC between the ranks' GPUs works): the in-kernel exchange over
      peer memory (``nsol_pd_plan_link_*``).  The boundary CTAs of every iteration store their new
      boundary planes straight into the neighbours' receive slots over NVLink and raise a flag; the
      neighbours' next launch waits on it.  ``iterate(n)`` queues n launches, nothing else.
    * ``halo="nccl"``: grouped NCCL send/recv between neighbours before every launch
      (``HaloExchanger``), optionally overlapped with the interior chunks (``overlap=True``).
    ``halo="auto"`` tries p2p and falls back to nccl (all ranks together, with a warning)."""

    def __init__(self, ctx, desc, plane_numel, np_dtype, rank, world, device, halo="auto"):
        import ctypes as C
        self.C = C
        self.ctx = ctx
        self.np_dtype = np_dtype
        self.plane_numel = plane_numel
        self.device = device
        self.rank, self.world = rank, world
        h = C.c_void_p()
        ctx.check(ctx.lib.nsol_pd_plan_create(ctx.handle, C.byref(desc), C.byref(h)))
        self.plan = h
        self._views = {}
        self.halo = None
        self.mode = "single" if world == 1 else halo
        if world == 1:
            return
        if self.mode not in ("auto", "p2p", "nccl"):
            raise ValueError("halo must be 'auto', 'p2p' or 'nccl'")
        if self.mode in ("auto", "p2p"):
            err = self._connect_link()
            if err is None:
                self.mode = "p2p"
            elif self.mode == "p2p":
                raise RuntimeError("in-kernel halo exchange unavailable: %s" % err)
            else:
                if rank == 0:
                    import sys
                    sys.stderr.write("nsol_b200: peer-memory halo exchange unavailable (%s); using NCCL send/recv\n" % err)
                self.mode = "nccl"
        if self.mode == "nccl":
            import torch
            tdtype = torch.float32 if np.dtype(np_dtype) == np.float32 else torch.float64
            self.halo = HaloExchanger(rank, world, plane_numel, tdtype, device)
            ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
            ctx.check(ctx.lib.nsol_pd_plan_set_halo(self.plan, ptr(self.halo.xbar_above), ptr(self.halo.xbar_below),
                                                   ptr(self.halo.pz_below)))

    def _connect_link(self):
        """Exchange the CUDA IPC handles of the link blocks with the neighbours and map theirs.
        Returns None on success (on every rank) or a message (on every rank)."""
        import torch.distributed as dist
        C, lib, ctx = self.C, self.ctx.lib, self.ctx
        handle = (C.c_char * 64)()
        err = None
        try:
            ctx.check(lib.nsol_pd_plan_link_ipc_handle(self.plan, handle))
        except Exception as e:      # e.g. plane not a multiple of 16 bytes
            err = "rank %d: %s" % (self.rank, e)
        mine = bytes(handle.raw) if err is None else None
        handles = [None] * self.world
        dist.all_gather_object(handles, mine)
        if any(hd is None for hd in handles):
            err = err or "a rank could not export its link block"
        else:
            below = handles[self.rank - 1] if self.rank > 0 else None
            above = handles[self.rank + 1] if self.rank < self.world - 1 else None
            try:
                ctx.check(lib.nsol_pd_plan_link_open(self.plan, below, above))
            except Exception as e:
                err = "rank %d: %s" % (self.rank, e)
        errs = [None] * self.world
        dist.all_gather_object(errs, err)
        errs = [e for e in errs if e]
        if errs and err is None:
            # somebody failed: this rank must not stay in link mode on its own
            raise RuntimeError("in-kernel halo exchange: inconsistent setup (%s)" % errs[0])
        return errs[0] if errs else None

    def _boundary_tensors(self, upcoming=False):
        """Boundary planes of the current state, or (upcoming=True) of the state the running
        split iteration is writing."""
        C = self.C
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        fn = self.ctx.lib.nsol_pd_plan_boundary_planes_next if upcoming else self.ctx.lib.nsol_pd_plan_boundary_planes
        self.ctx.check(fn(self.plan, C.byref(a), C.byref(b), C.byref(c)))
        key = (a.value, b.value, c.value)
        if key not in self._views:
            self._views[key] = tuple(tensor_from_ptr(p, self.plane_numel, self.np_dtype, self.device) for p in key)
        return self._views[key]

    def iterate(self, n, stream, overlap=False):
        """n iterations (every rank must call this with the same n).
        p2p: n back-to-back launches, the halo exchange happens inside the kernels.
        nccl: exchange, then one full-iteration launch; with overlap (needs >= 3 z-chunks) every
        iteration is split: the two boundary chunks run first, their new boundary planes are
        exchanged while the interior chunks compute."""
        lib, ctx, plan = self.ctx.lib, self.ctx, self.plan
        if self.mode in ("single", "p2p"):
            ctx.check(lib.nsol_pd_plan_iterate(plan, n, stream))
            return
        if not overlap or lib.nsol_pd_plan_chunks(plan) < 3:
            for _ in range(n):
                self.halo.exchange(*self._boundary_tensors())
                ctx.check(lib.nsol_pd_plan_iterate(plan, 1, stream))
            self._halo_fresh = False
            return
        if not getattr(self, "_halo_fresh", False):
            self.halo.exchange(*self._boundary_tensors())          # halos of the current state
        for _ in range(n):
            ctx.check(lib.nsol_pd_plan_iterate_part(plan, 1, stream))          # boundary chunks
            reqs = self.halo.start_exchange(*self._boundary_tensors(upcoming=True))
            ctx.check(lib.nsol_pd_plan_iterate_part(plan, 2, stream))          # interior chunks + advance
            self.halo.finish_exchange(reqs)
        self._halo_fresh = True     # the halos now belong to the current state

    def check(self, stream):
        """Synchronise and raise if an in-kernel halo wait timed out."""
        self.ctx.check(self.ctx.lib.nsol_pd_plan_link_status(self.plan, stream))

    def reset_host(self, b_host_ptr, x0_host_ptr, stream):
        self._halo_fresh = False
        self.ctx.check(self.ctx.lib.nsol_pd_plan_reset_host(self.plan, b_host_ptr, x0_host_ptr, stream))

    def reset_dev(self, b_dev_ptr, x0_dev_ptr, stream):
        self._halo_fresh = False
        self.ctx.check(self.ctx.lib.nsol_pd_plan_reset_dev(self.plan, b_dev_ptr, x0_dev_ptr, stream))

    def close(self):
        if self.plan is not None:
            if self.mode == "p2p":
                # the neighbours' kernels write into this rank's link block: everybody finishes first
                import torch
                import torch.distributed as dist
                torch.cuda.synchronize()
                dist.barrier()
            self.ctx.lib.nsol_pd_plan_destroy(self.plan)
            self.plan = None
