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


def _peer(dist, group, r):
    """Rank ``r`` of ``group`` as the global rank the P2P ops address."""
    return r if group is None else dist.get_global_rank(group, r)


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
            up = _peer(dist, self.group, self.rank + 1)
            ops.append(dist.P2POp(dist.isend, xbar_last, up, self.group))
            ops.append(dist.P2POp(dist.isend, pz_last, up, self.group))
            ops.append(dist.P2POp(dist.irecv, self.xbar_above, up, self.group))
        if self.has_below:
            dn = _peer(dist, self.group, self.rank - 1)
            ops.append(dist.P2POp(dist.isend, xbar_first, dn, self.group))
            ops.append(dist.P2POp(dist.irecv, self.xbar_below, dn, self.group))
            ops.append(dist.P2POp(dist.irecv, self.pz_below, dn, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish_exchange(reqs):
        for req in reqs:
            req.wait()      # NCCL: makes the current stream wait; gloo: blocks the host

    def exchange(self, xbar_first, xbar_last, pz_last):
        self.finish_exchange(self.start_exchange(xbar_first, xbar_last, pz_last))


def pipelined_schedule(groups, iterations, depth):
    """The order in which ``nsol_pd_plan_solve_host`` (csrc/pd_kernels.cu, pd_solve_pipelined) queues its work for a slab cut
    into ``groups`` transfer groups, as a list of steps -- a host-side mirror of the C++ loop, used by the CPU tests to check
    the wavefront order and the halo exchange by iteration number without a GPU.  Groups are numbered in ARRIVAL order
    (spatial group ``c`` when the transfers run bottom-up, ``groups - 1 - c`` top-down).  Steps:

    * ``("reset", c)``: group c has landed, its start state exists (a boundary group publishes it to its neighbour);
    * ``("advance", c0, c1, it)``: iteration ``it`` (0-based) on the arrival groups c0 ... c1, one launch;
    * ``("full", it)``: iteration ``it`` on the whole slab;
    * ``("download", c)``: group c holds the final state and goes back to the host.
    """
    d_up = min(depth, iterations // 2)
    d_dn = min(depth, iterations - d_up)
    done = [-1] * groups
    steps = []
    for j in range(groups + d_up):
        if j < groups:
            steps.append(("reset", j))
            done[j] = 0
        for c in range(min(j - 1, groups - 1), max(0, j - d_up) - 1, -1):
            steps.append(("advance", c, c, done[c]))
            done[c] += 1
    for it in range(d_up, iterations - d_dn):
        steps.append(("full", it))
    done = [iterations - d_dn] * groups
    for t in range(1, d_dn + 1):
        top = min(d_dn - t, groups - 1)
        steps.append(("advance", 0, top, iterations - d_dn + t - 1))
        for c in range(top + 1):
            done[c] += 1
    for j in range(groups):
        for c in range(min(j + d_dn - 1, groups - 1), j - 1, -1):
            if done[c] < iterations and (c + 1 >= groups or done[c + 1] >= done[c]):
                steps.append(("advance", c, c, done[c]))
                done[c] += 1
        assert done[j] == iterations
        steps.append(("download", j))
    return steps


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

    def __init__(self, ctx, desc, plane_numel, np_dtype, rank, world, device, halo="auto", group=None):
        import ctypes as C
        self.C = C
        self.ctx = ctx
        self.group = group
        self._unchecked = False      # launches queued since the last link-status check
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
                # pipelined host solves (nsol_pd_plan_solve_host): neighbouring slabs move their transfer groups in opposite
                # directions, so the wavefront of iterations continues through the slab boundaries
                ctx.check(ctx.lib.nsol_pd_plan_set_pipe_direction(self.plan, -1 if rank % 2 else 1))
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
            self.halo = HaloExchanger(rank, world, plane_numel, tdtype, device, group)
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
        dist.all_gather_object(handles, mine, group=self.group)
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
        dist.all_gather_object(errs, err, group=self.group)
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
            self._unchecked = self._unchecked or self.mode == "p2p"
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
        """Synchronise and raise if an in-kernel halo wait timed out (the iterates are then invalid: the
        kernels carried on with stale halo planes)."""
        self._unchecked = False
        self.ctx.check(self.ctx.lib.nsol_pd_plan_link_status(self.plan, stream))

    def reset_host(self, b_host_ptr, x0_host_ptr, stream, check=True):
        """New solve.  With the in-kernel exchange, ``check`` first verifies that no halo wait of the previous
        solve timed out (synchronises; pass False inside a timed loop and call ``check()`` afterwards)."""
        if check and self._unchecked:
            self.check(stream)
        self._halo_fresh = False
        self.ctx.check(self.ctx.lib.nsol_pd_plan_reset_host(self.plan, b_host_ptr, x0_host_ptr, stream))

    def solve_host(self, b_host_ptr, x0_host_ptr, iterations, x_host_ptr, stream):
        """One whole solve from / to host memory (every rank, same ``iterations``): upload, iterations and download of the slab
        overlap where the library can pipeline them (single rank or in-kernel halo exchange, page-locked buffers); with
        NCCL halos it is reset + iterate + download.  Synchronous; raises if a halo wait timed out."""
        lib, ctx, plan = self.ctx.lib, self.ctx, self.plan
        if self._unchecked:
            self.check(stream)
        if self.mode in ("single", "p2p"):
            ctx.check(lib.nsol_pd_plan_solve_host(plan, b_host_ptr, x0_host_ptr, iterations, x_host_ptr, stream))
            return
        self.reset_host(b_host_ptr, x0_host_ptr, stream)
        self.iterate(iterations, stream)
        ctx.check(lib.nsol_pd_plan_get_x_host(plan, x_host_ptr, stream))

    def reset_dev(self, b_dev_ptr, x0_dev_ptr, stream, check=True):
        if check and self._unchecked:
            self.check(stream)
        self._halo_fresh = False
        self.ctx.check(self.ctx.lib.nsol_pd_plan_reset_dev(self.plan, b_dev_ptr, x0_dev_ptr, stream))

    def close(self):
        if self.plan is not None:
            err = None
            if self.mode == "p2p":
                # the neighbours' kernels write into this rank's link block: everybody finishes first
                import torch
                import torch.distributed as dist
                torch.cuda.synchronize()
                if self._unchecked:
                    try:
                        self.check(None)
                    except RuntimeError as e:       # still tear down in lockstep with the other ranks
                        err = e
                dist.barrier(group=self.group)
            self.ctx.lib.nsol_pd_plan_destroy(self.plan)
            self.plan = None
            if err is not None:
                raise err


# ---------------------------------------------------------------------------------------------
# z-slab decomposition of the ADMM / LSMR path
# ---------------------------------------------------------------------------------------------
def slab_admm_program(iterations, iter_max, alpha, rho, lo=0.0, hi=float("inf")):
    """The ADMM TV-L2 run (nsol/admm_linear_solver.py:165-253; inner solve nsol/tikhonov_linear_solver.py:
    226-274 + scipy lsmr.py:239-479) as a flat list of steps for one slab:
        ("exchange", [buffer kinds])      neighbour exchange of halo planes (before the next phase)
        ("phase", id, p0, p1, i0)         nsol_lsmr_slab_phase
        ("allreduce",)                    sum of the one-double reduction buffer over all slabs
    Every rank executes the same list in lockstep; the device-side ``done`` flag of LSMR turns the
    remaining phases of a solve into no-ops, so the list never depends on data."""
    from nsol_b200 import _lib as L
    sa = float(np.sqrt(rho))
    steps = [("exchange", [L.SLAB_X]), ("phase", L.PH_ADMM_INIT, 0.0, 0.0, 0)]
    for _ in range(int(iterations)):
        steps += [("phase", L.PH_RHS, sa, 0.0, 1), ("allreduce",), ("phase", L.PH_SCAL_INIT_BETA, sa, 0.0, int(iter_max)),
                  ("exchange", [L.SLAB_U0, L.SLAB_UZ]), ("phase", L.PH_ADJ_FIRST, 0.0, 0.0, 0), ("allreduce",),
                  ("phase", L.PH_SCAL_INIT_ALPHA, 0.0, 0.0, 0)]
        for _ in range(int(iter_max)):
            steps += [("exchange", [L.SLAB_V]), ("phase", L.PH_FWD, 0.0, 0.0, 0), ("allreduce",), ("phase", L.PH_SCAL_BETA, 0.0, 0.0, 0),
                      ("exchange", [L.SLAB_U0, L.SLAB_UZ]), ("phase", L.PH_ADJ, 0.0, 0.0, 0), ("allreduce",),
                      ("phase", L.PH_SCAL_ALPHA, 0.0, 0.0, 0),
                      ("phase", L.PH_UPDATE, 0.0, 0.0, 0), ("allreduce",), ("phase", L.PH_SCAL_TESTS, 0.0, 0.0, 0)]
        steps += [("phase", L.PH_CLIP, float(lo), float(hi), 0),
                  ("exchange", [L.SLAB_X]), ("phase", L.PH_ADMM_SHRINK, float(alpha) / float(rho), 0.0, 0)]
    return steps


class SlabLsq(object):
    """One rank's slab of the stacked least-squares problem: an ``nsol_lsmr_plan`` in slab mode plus
    views of its exchange buffers."""

    def __init__(self, ctx, info, dtype, rank, world):
        """info: {"shape": local slab shape, "spacing", "a_kind": "conv" | "identity", "a_op" (taps), "b_kind": "grad"}"""
        import ctypes as C
        from nsol_b200 import _lib as L
        from nsol_b200.linear_solver import LsmrPlan
        self.C, self.L, self.ctx = C, L, ctx
        self.rank, self.world = rank, world
        self.plan = LsmrPlan(info, dtype)
        self.handle = self.plan.handle
        self.dcode = L.dtype_code(dtype)
        self.np_dtype = L.np_dtype(self.dcode)
        self.shape = tuple(int(s) for s in info["shape"])
        self.n = int(np.prod(self.shape))
        self.plane = self.n // self.shape[0]
        ctx.check(ctx.lib.nsol_lsmr_plan_slab(self.handle, 1 if rank > 0 else 0, 1 if rank < world - 1 else 0))
        b, x, ss = C.c_void_p(), C.c_void_p(), C.c_void_p()
        ctx.check(ctx.lib.nsol_lsmr_slab_arrays(self.handle, C.byref(b), C.byref(x), C.byref(ss)))
        self.b_ptr, self.x_ptr, self.ss_ptr = b, x, ss
        self.buffers = {}
        for kind in (L.SLAB_V, L.SLAB_U0, L.SLAB_UZ, L.SLAB_X):
            sf, sl, rl, rh, npl = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
            ctx.check(ctx.lib.nsol_lsmr_slab_buffers(self.handle, kind, C.byref(sf), C.byref(sl), C.byref(rl), C.byref(rh), C.byref(npl)))
            self.buffers[kind] = {"send_first": sf.value, "send_last": sl.value, "recv_lo": rl.value, "recv_hi": rh.value,
                                  "numel": npl.value * self.plane}

    def upload(self, b_scaled, x0_scaled, stream=None):
        """b and the start value, float64 host arrays already in solver units (divided by x_scale)."""
        ctx, L = self.ctx, self.L
        for ptr, arr in ((self.b_ptr, b_scaled), (self.x_ptr, x0_scaled)):
            arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
            if arr.size != self.n:
                raise ValueError("slab array has %d values, the slab %d" % (arr.size, self.n))
            stage = ctx.device_alloc(arr.nbytes).upload(arr, stream)
            ctx.check(ctx.lib.nsol_scale_convert(ctx.handle, self.n, L.F64, stage.ptr, self.dcode, ptr, 1.0, 0, stream))
            ctx.sync(stream)
            stage.free()

    def download(self, x_scale=1.0, stream=None):
        ctx, L = self.ctx, self.L
        stage = ctx.device_alloc(self.n * 8)
        ctx.check(ctx.lib.nsol_scale_convert(ctx.handle, self.n, self.dcode, self.x_ptr, L.F64, stage.ptr, float(x_scale), 0, stream))
        out = stage.download((self.n,), np.float64, stream)
        stage.free()
        return out

    def phase(self, pid, p0, p1, i0, stream=None):
        self.ctx.check(self.ctx.lib.nsol_lsmr_slab_phase(self.handle, pid, p0, p1, i0, stream))

    def close(self):
        self.plan.close()


def run_slab_admm_emulated(slabs, steps, stream=None):
    """Execute a step program for ALL slabs inside one process on one GPU (tests): the exchange is a
    set of device copies, the all-reduce a host-side sum in rank order."""
    import ctypes as C
    ctx = slabs[0].ctx
    L = slabs[0].L
    world = len(slabs)
    esz = 4 if slabs[0].dcode == L.F32 else 8
    for st in steps:
        if st[0] == "phase":
            for sl in slabs:
                sl.phase(st[1], st[2], st[3], st[4], stream)
        elif st[0] == "exchange":
            for kind in st[1]:
                periodic = kind in (L.SLAB_V, L.SLAB_U0)
                for r, sl in enumerate(slabs):
                    buf = sl.buffers[kind]
                    nbytes = buf["numel"] * esz
                    up, dn = r + 1, r - 1
                    if periodic:
                        up, dn = up % world, dn % world
                    if buf["send_last"] and 0 <= up < world and slabs[up].buffers[kind]["recv_lo"]:
                        ctx.check(ctx.lib.nsol_memcpy_d2d(ctx.handle, slabs[up].buffers[kind]["recv_lo"], buf["send_last"], nbytes, stream))
                    if buf["send_first"] and 0 <= dn < world and slabs[dn].buffers[kind]["recv_hi"]:
                        ctx.check(ctx.lib.nsol_memcpy_d2d(ctx.handle, slabs[dn].buffers[kind]["recv_hi"], buf["send_first"], nbytes, stream))
        else:
            vals = np.zeros(world)
            for r, sl in enumerate(slabs):
                one = np.zeros(1)
                ctx.check(ctx.lib.nsol_memcpy_d2h(ctx.handle, one.ctypes.data, sl.ss_ptr, 8, stream))
                ctx.sync(stream)
                vals[r] = one[0]
            total = np.array([np.sum(vals)])      # rank order, like a tree-less all-reduce
            for sl in slabs:
                ctx.check(ctx.lib.nsol_memcpy_h2d(ctx.handle, sl.ss_ptr, total.ctypes.data, 8, stream))
                ctx.sync(stream)


def exchange_slab_halos(dist, group, rank, world, items):
    """One grouped neighbour exchange.  items: [(views, periodic)], views = dict of 1-D tensors
    (or None) "send_first", "send_last" (this rank's first / last planes), "recv_lo" (<- the lower
    neighbour's last planes), "recv_hi" (<- the upper neighbour's first planes).  periodic: the slabs
    form a ring (blur, mode="wrap"); otherwise the global ends have no neighbour (gradient)."""
    ops = []
    for v, periodic in items:
        up, dn = rank + 1, rank - 1
        if periodic:
            up, dn = up % world, dn % world
        has_up, has_dn = 0 <= up < world, 0 <= dn < world
        if world == 1:
            if periodic:       # the slab is its own ring neighbour
                v["recv_lo"].copy_(v["send_last"])
                v["recv_hi"].copy_(v["send_first"])
            continue
        # order matters when both neighbours are the same rank (world == 2, ring): the k-th send to a
        # peer pairs with the peer's k-th receive from this rank
        gup, gdn = (_peer(dist, group, up) if has_up else up), (_peer(dist, group, dn) if has_dn else dn)
        if v["send_first"] is not None and has_dn:
            ops.append(dist.P2POp(dist.isend, v["send_first"], gdn, group))
        if v["send_last"] is not None and has_up:
            ops.append(dist.P2POp(dist.isend, v["send_last"], gup, group))
        if v["recv_hi"] is not None and has_up and v["send_first"] is not None:
            ops.append(dist.P2POp(dist.irecv, v["recv_hi"], gup, group))
        if v["recv_lo"] is not None and has_dn and v["send_last"] is not None:
            ops.append(dist.P2POp(dist.irecv, v["recv_lo"], gdn, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class SlabADMM(object):
    """ADMM TV-L2 deconvolution of one tall volume, z-slab sharded over the ranks of a
    ``torch.distributed`` group (NCCL): kernel-radius halo planes of the blur travel on a ring
    (periodic boundary), one-plane halos of the gradient between neighbours, and three scalar
    all-reduces per LSMR iteration (||u||^2, ||v||^2, ||x||^2) -- SURVEY.md 8e.  All ranks call ``run``
    together."""

    def __init__(self, ctx, info, dtype, rank, world, device, group=None):
        import torch
        self.torch = torch
        self.group = group
        self.device = device
        self.slab = SlabLsq(ctx, info, dtype, rank, world)
        sl = self.slab
        self.ss = tensor_from_ptr(sl.ss_ptr.value, 1, np.float64, device)
        self.views = {}
        for kind, buf in sl.buffers.items():
            self.views[kind] = {k: (tensor_from_ptr(buf[k], buf["numel"], sl.np_dtype, device) if buf[k] else None)
                                for k in ("send_first", "send_last", "recv_lo", "recv_hi")}

    def _exchange(self, kinds):
        L = self.slab.L
        exchange_slab_halos(self.torch.distributed, self.group, self.slab.rank, self.slab.world,
                            [(self.views[k], k in (L.SLAB_V, L.SLAB_U0)) for k in kinds])

    def _execute(self, st, stream):
        if st[0] == "phase":
            self.slab.phase(st[1], st[2], st[3], st[4], stream)
        elif st[0] == "exchange":
            self._exchange(st[1])
        elif self.slab.world > 1:
            self.torch.distributed.all_reduce(self.ss, group=self.group)

    def run(self, b_scaled, x0_scaled, alpha, rho, iterations, iter_max, stream=None, graph=True):
        """b, x0: this rank's slab in solver units (float64 host arrays).  Returns the slab of the result
        (solver units, float64).

        The step program of ONE outer iteration (~60 kernels, 21 grouped send/recv rounds and 32 one-double all-reduces
        for iter_max = 10) does not depend on data -- LSMR stops through device-side flags -- so with ``graph`` (default
        on NCCL) it is captured once into a CUDA graph, NCCL calls included, and replayed ``iterations`` times: no host
        work per inner iteration (round 1 issued every exchange and all-reduce from Python, ~0.5-2 ms per inner iteration)."""
        torch = self.torch
        self.slab.upload(b_scaled, x0_scaled, stream)
        steps = slab_admm_program(iterations, iter_max, alpha, rho)
        head, body = steps[:2], steps[2:]            # [exchange X, ADMM_INIT] + iterations x (one outer iteration)
        per_outer = len(body) // iterations if iterations > 0 else 0
        use_graph = (graph and self.slab.world > 1 and iterations >= 2 and stream is None
                     and torch.distributed.get_backend(self.group) == "nccl")
        timed = stream is None and torch.cuda.is_available()
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for st in head:
            self._execute(st, stream)
        if use_graph:
            import ctypes as C
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            # thread_local: NCCL's watchdog thread polls events while this thread captures
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                cs = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                for st in body[:per_outer]:
                    self._execute(st, cs)
            torch.cuda.current_stream().wait_stream(side)
            if timed:
                e0.record()          # the capture is set-up, not solve time
            for _ in range(iterations):
                g.replay()
            if timed:
                e1.record()
            torch.cuda.current_stream().synchronize()
            del g
        else:
            for st in body:
                self._execute(st, stream)
            if timed:
                e1.record()
        if timed:
            torch.cuda.current_stream().synchronize()
            self.last_device_ms = e0.elapsed_time(e1)       # device time of the ADMM iterations (CUDA events)
        self.last_used_graph = bool(use_graph)
        return self.slab.download(1.0, stream)

    def close(self):
        self.slab.close()
