"""Multi-process host logic on CPU (gloo, world_size 2 and 3): z-slab bounds, the 3-plane
halo exchange protocol of nsol_b200/distributed.py and the round-robin partition of sweep
points.  The local compute is a numpy slab iteration built from the oracle (test
infrastructure); the sharded run must reproduce the unsharded oracle bit for bit."""
import os
import socket

import numpy as np
import pytest

from oracle import nsol_oracle as orc
from nsol_b200.distributed import HaloExchanger, exchange_slab_halos, partition_round_robin, pipelined_schedule, slab_bounds


def test_slab_bounds_and_partition():
    for nz, world in ((512, 8), (10, 3), (5, 5), (7, 2)):
        spans = [slab_bounds(nz, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == nz
        for a, b in zip(spans[:-1], spans[1:]):
            assert a[1] == b[0]
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    items = sorted(i for r in range(8) for i in partition_round_robin(192, r, 8))
    assert items == list(range(192))
    assert len(partition_round_robin(192, 3, 8)) == 24     # BASELINE config 5: 24 runs per GPU


def slab_iteration(x, xbar, p, b, sched_row, xbar_above, xbar_below, pz_below, data="L2"):
    """One primal-dual iteration on a z-slab given the neighbours' boundary planes (or None at
    the global ends).  Same arithmetic as oracle.primal_dual_denoise, unit spacing, TV."""
    sigma, tau, tl, theta = sched_row
    nzl = x.shape[0]
    top = np.zeros_like(xbar[:1]) if xbar_above is None else xbar_above[None]
    ext = np.concatenate([xbar, top])                       # planes z_lo .. z_hi
    g = orc.grad(ext)
    n_ext = ext.shape[0]
    gx, gy, gz = g[:n_ext][:nzl], g[n_ext:2 * n_ext][:nzl], g[2 * n_ext:][:nzl]
    pn = [orc.prox_tv_conj(p[k] + sigma * gk, sigma) for k, gk in enumerate((gx, gy, gz))]
    if pz_below is None:
        pz_m = np.zeros_like(xbar[:1])
    else:
        pz_m = orc.prox_tv_conj(pz_below[None] + sigma * (1.0 * xbar[:1] + (-1.0) * xbar_below[None]), sigma)
    div = orc.forward_difference_adj(pn[0], 0) + orc.forward_difference_adj(pn[1], 1)
    div = div + orc.forward_difference_adj(np.concatenate([pz_m, pn[2]]), 2)[1:]
    y = x - tau * div
    xn = (y + tl * b) / (1.0 + tl)
    xbn = xn + theta * (xn - x)
    return xn, xbn, pn


def _worker(rank, world, port, shape, iterations, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(7)
    vol = rng.rand(*shape) * 255
    xs = float(vol.max())
    z_lo, z_hi = slab_bounds(shape[0], rank, world)
    b = vol[z_lo:z_hi] / xs
    x = b.copy()
    xbar = b.copy()
    p = [np.zeros_like(b) for _ in range(3)]
    sched = orc.pd_schedule("ALG2", 8.0, 0.05, iterations)
    plane = shape[1] * shape[2]
    halo = HaloExchanger(rank, world, plane, torch.float64, torch.device("cpu"))
    as_t = lambda a: torch.from_numpy(np.ascontiguousarray(a).reshape(-1))
    for it in range(iterations):
        halo.exchange(as_t(xbar[0]), as_t(xbar[-1]), as_t(p[2][-1]))
        above = halo.xbar_above.numpy().reshape(shape[1:]) if halo.has_above else None
        below = halo.xbar_below.numpy().reshape(shape[1:]) if halo.has_below else None
        pzb = halo.pz_below.numpy().reshape(shape[1:]) if halo.has_below else None
        x, xbar, p = slab_iteration(x, xbar, p, b, sched[it], above, below, pzb)
    gathered = [None] * world
    dist.all_gather_object(gathered, (z_lo, x * xs))
    if rank == 0:
        full = np.concatenate([g[1] for g in sorted(gathered, key=lambda t: t[0])])
        np.save(out, full)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 3])
def test_zslab_halo_exchange_reproduces_unsharded_oracle(tmp_path, world):
    import torch.multiprocessing as mp
    shape, iterations = (11, 6, 7), 6
    out = str(tmp_path / "sharded.npy")
    mp.spawn(_worker, args=(world, _free_port(), shape, iterations, out), nprocs=world, join=True)
    rng = np.random.RandomState(7)
    vol = rng.rand(*shape) * 255
    ref = orc.primal_dual_denoise(vol.reshape(-1), shape, reg="TV", data="L2", alpha=0.05, L2=8.0,
                                  iterations=iterations, x_scale=float(vol.max()))
    assert np.array_equal(np.load(out).reshape(-1), ref)


def _study_worker(rank, world, port, directory):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_parameter_study import _make
    solver, study = _make(directory)
    study.run()
    # every rank ran only its round-robin share of the 6 points
    runs = [c for c in solver.calls if c != "set_x0"]
    assert len(runs) == len(partition_round_robin(6, rank, world))
    dist.destroy_process_group()


def test_parameter_study_fans_out_over_ranks(tmp_path):
    """BASELINE config 5: sweep points are independent -> dealt round-robin to the ranks, no
    data-path communication; rank 0 gathers and writes the same files as a single process."""
    import torch.multiprocessing as mp
    from nsol_b200.reader_parameter_study import ReaderParameterStudy
    from test_parameter_study import _make
    multi = tmp_path / "multi"
    single = tmp_path / "single"
    multi.mkdir()
    single.mkdir()
    mp.spawn(_study_worker, args=(2, _free_port(), str(multi)), nprocs=2, join=True)
    _, study = _make(str(single))
    study.run()
    a = ReaderParameterStudy(str(multi), "Stub")
    b = ReaderParameterStudy(str(single), "Stub")
    a.read_study()
    b.read_study()
    assert a.get_parameters_to_line() == b.get_parameters_to_line()
    assert np.array_equal(a.get_results("SUM"), b.get_results("SUM"))
    ra, rb = a.get_reconstructions(), b.get_reconstructions()
    assert sorted(ra.files) == sorted(rb.files)
    for k in ra.files:
        assert np.array_equal(ra[k], rb[k])


def _halo_worker(rank, world, port, ghost, plane):
    """ADMM / LSMR slab exchange (SlabADMM._exchange): every rank fills tensors whose values encode
    (owner rank, plane, element) and checks what arrives in its receive buffers."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nz = ghost + 2

    def planes(owner, kind, lo, hi):
        z = np.arange(lo, hi)[:, None]
        e = np.arange(plane)[None, :]
        return (1000.0 * kind + 100.0 * owner + z + e / 1000.0).reshape(-1)

    def views(kind, count, first, last, lo, hi):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        return {"send_first": t(planes(rank, kind, 0, count)) if first else None,
                "send_last": t(planes(rank, kind, nz - count, nz)) if last else None,
                "recv_lo": torch.full((count * plane,), -1.0, dtype=torch.float64) if lo else None,
                "recv_hi": torch.full((count * plane,), -1.0, dtype=torch.float64) if hi else None}

    v = views(0, ghost, True, True, True, True)      # vhat: both ways, ring
    u0 = views(1, ghost, True, True, True, True)     # u block 0: both ways, ring
    uz = views(2, 1, False, True, True, False)       # u_z: last plane -> upper neighbour
    x = views(3, 1, True, False, False, True)        # x: first plane -> lower neighbour
    exchange_slab_halos(dist, None, rank, world, [(v, True)])
    exchange_slab_halos(dist, None, rank, world, [(u0, True), (uz, False)])
    exchange_slab_halos(dist, None, rank, world, [(x, False)])
    up, dn = (rank + 1) % world, (rank - 1) % world
    for kind, vw in ((0, v), (1, u0)):
        assert np.array_equal(vw["recv_lo"].numpy(), planes(dn, kind, nz - ghost, nz)), (rank, kind, "lo")
        assert np.array_equal(vw["recv_hi"].numpy(), planes(up, kind, 0, ghost)), (rank, kind, "hi")
    if rank > 0:
        assert np.array_equal(uz["recv_lo"].numpy(), planes(rank - 1, 2, nz - 1, nz))
    else:
        assert np.all(uz["recv_lo"].numpy() == -1.0)          # global bottom: untouched (zero boundary in the kernel)
    if rank < world - 1:
        assert np.array_equal(x["recv_hi"].numpy(), planes(rank + 1, 3, 0, 1))
    else:
        assert np.all(x["recv_hi"].numpy() == -1.0)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_admm_slab_halo_exchange_protocol(world):
    """Ring (periodic blur) and neighbour (gradient) exchanges of the z-slab ADMM driver deliver the
    right planes, including world == 2 where both ring neighbours are the same rank."""
    import torch.multiprocessing as mp
    mp.spawn(_halo_worker, args=(world, _free_port(), 3, 5), nprocs=world, join=True)


def test_slab_admm_program_shape():
    from nsol_b200 import _lib as L
    from nsol_b200.distributed import slab_admm_program
    steps = slab_admm_program(iterations=2, iter_max=3, alpha=0.01, rho=0.1)
    assert steps[0] == ("exchange", [L.SLAB_X]) and steps[1][1] == L.PH_ADMM_INIT
    assert sum(1 for s in steps if s[0] == "allreduce") == 2 * (2 + 3 * 3)
    phases = [s[1] for s in steps if s[0] == "phase"]
    assert phases.count(L.PH_FWD) == 6 and phases.count(L.PH_ADMM_SHRINK) == 2 and phases.count(L.PH_CLIP) == 2
    # every FWD is preceded by an exchange of vhat, every ADJ by an exchange of both u halos
    for i, s in enumerate(steps):
        if s[0] == "phase" and s[1] == L.PH_FWD:
            assert steps[i - 1] == ("exchange", [L.SLAB_V])
        if s[0] == "phase" and s[1] in (L.PH_ADJ, L.PH_ADJ_FIRST):
            assert steps[i - 1] == ("exchange", [L.SLAB_U0, L.SLAB_UZ])
        if s[0] == "phase" and s[1] == L.PH_ADMM_SHRINK:
            assert abs(s[2] - 0.1) < 1e-15


def _admm_slab_worker(rank, world, port, shape, var, params, out):
    """ADMM TV-L2 on z-slabs with gloo: the oracle's ADMM / LSMR recurrences run on the local slab of every
    rank; the operators fetch their halo planes with exchange_slab_halos exactly like SlabADMM (ring for the
    periodic blur, neighbours for grad / grad_adj) and the norms are all-reduced."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dim = len(shape)
    rng = np.random.RandomState(41)
    vol = rng.rand(*shape) * 200 + 20
    z_lo, z_hi = slab_bounds(shape[0], rank, world)
    loc = (z_hi - z_lo,) + tuple(shape[1:])
    plane = int(np.prod(shape[1:]))
    taps = orc.separable_taps(orc.gaussian_kernel(dim, np.diag(var)))
    r = (taps[0].size - 1) // 2
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).reshape(-1))
    buf = lambda planes: torch.zeros(planes * plane, dtype=torch.float64)

    def blur(x_flat):
        x = x_flat.reshape(loc)
        views = {"send_first": t(x[:r]), "send_last": t(x[-r:]), "recv_lo": buf(r), "recv_hi": buf(r)}
        exchange_slab_halos(dist, None, rank, world, [(views, True)])            # ring: periodic boundary
        ext = np.concatenate([views["recv_lo"].numpy().reshape((r,) + loc[1:]), x,
                              views["recv_hi"].numpy().reshape((r,) + loc[1:])])
        acc = np.zeros(loc)
        for k in range(taps[0].size):                                             # same tap order as the oracle's axis-0 pass
            acc += taps[0][k] * ext[2 * r - k: 2 * r - k + loc[0]]
        return orc.convolve_wrap_separable(acc, [np.ones(1)] + list(taps[1:])).reshape(-1)

    def grad(x_flat):
        x = x_flat.reshape(loc)
        views = {"send_first": t(x[0]), "send_last": None, "recv_lo": None, "recv_hi": buf(1)}
        exchange_slab_halos(dist, None, rank, world, [(views, False)])           # zero boundary at the global top
        top = views["recv_hi"].numpy().reshape((1,) + loc[1:]) if rank < world - 1 else np.zeros((1,) + loc[1:])
        g = orc.grad(np.concatenate([x, top]))
        n_ext = loc[0] + 1
        return np.concatenate([g[k * n_ext:(k + 1) * n_ext][:loc[0]] for k in range(dim)]).reshape(-1)

    def grad_adj(p_flat):
        p = p_flat.reshape((dim,) + loc)
        pz = p[dim - 1]
        views = {"send_first": None, "send_last": t(pz[-1]), "recv_lo": buf(1), "recv_hi": None}
        exchange_slab_halos(dist, None, rank, world, [(views, False)])
        below = views["recv_lo"].numpy().reshape((1,) + loc[1:]) if rank > 0 else np.zeros((1,) + loc[1:])
        out = orc.forward_difference_adj(p[0], 0)
        for k in range(1, dim - 1):
            out = out + orc.forward_difference_adj(p[k], k)
        out = out + orc.forward_difference_adj(np.concatenate([below, pz]), dim - 1)[1:]
        return out.reshape(-1)

    def norm(v):
        ss = torch.tensor([float(np.dot(v, v))], dtype=torch.float64)
        dist.all_reduce(ss)
        return np.float64(np.sqrt(ss.item()))

    mine = vol[z_lo:z_hi].reshape(-1)
    xs = 220.0
    x = orc.admm_tv(blur, blur, grad, grad_adj, mine, mine, dim, x_scale=xs, norm=norm, **params)
    gathered = [None] * world
    dist.all_gather_object(gathered, (z_lo, x))
    if rank == 0:
        np.save(out, np.concatenate([g[1] for g in sorted(gathered, key=lambda q: q[0])]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (16, 12)), (3, (13, 6, 8))])
def test_admm_zslab_with_gloo_matches_unsharded_oracle(tmp_path, world, shape):
    """SURVEY.md 8e on CPU: ring exchange of the blur halos (r = 3 planes, periodic), neighbour exchange of the
    gradient halos (zero boundary at the global ends) and all-reduced LSMR norms reproduce the unsharded ADMM."""
    import torch.multiprocessing as mp
    var = [1.0] * len(shape)
    params = dict(alpha=0.02, rho=0.2, iterations=3, iter_max=5)
    out = str(tmp_path / "admm_slabs.npy")
    mp.spawn(_admm_slab_worker, args=(world, _free_port(), shape, var, params, out), nprocs=world, join=True)
    rng = np.random.RandomState(41)
    vol = rng.rand(*shape) * 200 + 20
    A, A_adj, D, D_adj = orc.deconvolution_operators(shape, np.diag(var))
    ref = orc.admm_tv(A, A_adj, D, D_adj, vol.reshape(-1), vol.reshape(-1), len(shape), x_scale=220.0, **params)
    got = np.load(out)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-10


# ---------------------------------------------------------------------------------------------
# pipelined host solve through linked z-slabs: wavefront order + halo exchange by iteration number
# ---------------------------------------------------------------------------------------------
def _pipelined_worker(rank, world, port, shape, iterations, planes, depth, out):
    """Every rank executes the step list of nsol_pd_plan_solve_host (pipelined_schedule) on numpy arrays laid out like the
    plan -- two parity copies of xbar and p, x in place -- so a step that ran before its inputs were at the right iteration
    (a wavefront-order violation) reads the wrong state and the result differs from the unsharded oracle.  Odd ranks move
    their groups top-down.  Boundary planes travel as tagged messages (tag = iteration number of the state)."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(7)
    vol = rng.rand(*shape) * 255
    xs = float(vol.max())
    z_lo, z_hi = slab_bounds(shape[0], rank, world)
    nzl = z_hi - z_lo
    b = vol[z_lo:z_hi] / xs
    x = np.full_like(b, np.nan)                              # nothing is valid before its group has "landed"
    xbar = [np.full_like(b, np.nan), np.full_like(b, np.nan)]
    p = [[np.full_like(b, np.nan) for _ in range(3)] for _ in range(2)]
    sched = orc.pd_schedule("ALG2", 8.0, 0.05, iterations)
    ng = (nzl + planes - 1) // planes
    down = world > 1 and rank % 2 == 1
    spatial = lambda c: ng - 1 - c if down else c
    lo_of = lambda g: min(g * planes, nzl)
    has_below, has_above = rank > 0, rank < world - 1
    pending = []

    def send(plane, dst, kind, state):                       # kind 0: my first xbar plane, 1: my last xbar plane, 2: my last p_z plane
        pending.append(dist.isend(torch.from_numpy(np.ascontiguousarray(plane)), dst, tag=state * 4 + kind))

    def recv(src, kind, state):
        t = torch.empty(shape[1:], dtype=torch.float64)
        dist.recv(t, src, tag=state * 4 + kind)
        return t.numpy()

    def publish(plo, phi, state, par):
        if state == iterations:
            return                                           # nobody consumes the final state (on the GPU it just lands in the slots)
        if plo == 0 and has_below:
            send(xbar[par][0], rank - 1, 0, state)
        if phi == nzl and has_above:
            send(xbar[par][-1], rank + 1, 1, state)
            send(p[par][2][-1], rank + 1, 2, state)

    def advance(plo, phi, it):
        rd, wr = it & 1, (it + 1) & 1
        if phi < nzl:
            above = xbar[rd][phi]
        else:
            above = recv(rank + 1, 0, it) if has_above else None
        if plo > 0:
            below, pzb = xbar[rd][plo - 1], p[rd][2][plo - 1]
        elif has_below:
            below, pzb = recv(rank - 1, 1, it), recv(rank - 1, 2, it)
        else:
            below = pzb = None
        xn, xbn, pn = slab_iteration(x[plo:phi], xbar[rd][plo:phi], [q[plo:phi] for q in p[rd]], b[plo:phi], sched[it],
                                     above, below, pzb)
        assert np.all(np.isfinite(xn)), ("read a plane that was not valid yet", rank, plo, phi, it)
        x[plo:phi] = xn
        xbar[wr][plo:phi] = xbn
        for k in range(3):
            p[wr][k][plo:phi] = pn[k]
        publish(plo, phi, it + 1, wr)

    result = np.full_like(b, np.nan)
    for step in pipelined_schedule(ng, iterations, depth):
        if step[0] == "reset":
            g = spatial(step[1])
            plo, phi = lo_of(g), lo_of(g + 1)
            x[plo:phi] = b[plo:phi]
            xbar[0][plo:phi] = b[plo:phi]
            for k in range(3):
                p[0][k][plo:phi] = 0.0
            publish(plo, phi, 0, 0)
        elif step[0] == "advance":
            ga, gb = spatial(step[1]), spatial(step[2])
            advance(lo_of(min(ga, gb)), lo_of(max(ga, gb) + 1), step[3])
        elif step[0] == "full":
            advance(0, nzl, step[1])
        else:
            g = spatial(step[1])
            result[lo_of(g):lo_of(g + 1)] = x[lo_of(g):lo_of(g + 1)] * xs
    for r in pending:
        r.wait()
    gathered = [None] * world
    dist.all_gather_object(gathered, (z_lo, result))
    if rank == 0:
        np.save(out, np.concatenate([g[1] for g in sorted(gathered, key=lambda t: t[0])]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,iterations,planes,depth", [
    (1, (13, 4, 5), 9, 2, 3),          # one rank: the plain wavefront (upload / download overlap of nsol_pd_plan_solve_host)
    (2, (23, 4, 5), 9, 2, 3),          # two slabs, opposite directions, ragged last groups
    (3, (25, 4, 5), 8, 3, 10),         # three slabs, depth clipped by the iteration count
    (2, (9, 4, 5), 5, 16, 4),          # one group per slab: both boundaries in every launch
])
def test_pipelined_wavefront_through_linked_slabs(tmp_path, world, shape, iterations, planes, depth):
    import torch.multiprocessing as mp
    out = str(tmp_path / "piped.npy")
    mp.spawn(_pipelined_worker, args=(world, _free_port(), shape, iterations, planes, depth, out), nprocs=world, join=True)
    rng = np.random.RandomState(7)
    vol = rng.rand(*shape) * 255
    ref = orc.primal_dual_denoise(vol.reshape(-1), shape, reg="TV", data="L2", alpha=0.05, L2=8.0,
                                  iterations=iterations, x_scale=float(vol.max()))
    assert np.array_equal(np.load(out).reshape(-1), ref)


def test_pipelined_schedule_counts():
    """Every group is reset once, advanced `iterations` times in total (alone, merged or as part of a whole-slab step) and
    downloaded once, in arrival order."""
    for groups, iterations, depth in ((1, 0, 10), (1, 7, 10), (5, 1, 10), (9, 24, 10), (32, 100, 10), (4, 9, 2)):
        steps = pipelined_schedule(groups, iterations, depth)
        count = [0] * groups
        for st in steps:
            if st[0] == "advance":
                for c in range(st[1], st[2] + 1):
                    assert count[c] == st[3], (groups, iterations, depth, st)
                    count[c] += 1
            elif st[0] == "full":
                assert all(c == st[1] for c in count)
                count = [c + 1 for c in count]
        assert count == [iterations] * groups
        assert [st[1] for st in steps if st[0] == "reset"] == list(range(groups))
        assert [st[1] for st in steps if st[0] == "download"] == list(range(groups))


def test_pipelined_schedule_obeys_the_wavefront_rule():
    """Iteration `it` of a group reads state `it` of its two neighbours out of the parity buffers: when it is queued, both
    neighbours must have done at least `it` iterations (their state `it` exists) and at most `it + 1` (it has not been
    overwritten) -- for every step of every schedule, and merged launches must cover groups at one common iteration."""
    rng = np.random.RandomState(0)
    cases = [(g, n, d) for g in (1, 2, 3, 7, 16, 33) for n in (0, 1, 2, 5, 24, 100) for d in (1, 2, 10, 40)]
    cases += [tuple(int(v) for v in rng.randint(1, 60, size=3)) for _ in range(200)]
    for groups, iterations, depth in cases:
        done = [None] * groups                              # None: not on the device yet
        for st in pipelined_schedule(groups, iterations, depth):
            if st[0] == "reset":
                assert done[st[1]] is None
                done[st[1]] = 0
            elif st[0] in ("advance", "full"):
                c0, c1, it = (st[1], st[2], st[3]) if st[0] == "advance" else (0, groups - 1, st[1])
                for c in range(c0, c1 + 1):
                    assert done[c] == it, (groups, iterations, depth, st)
                for nb in (c0 - 1, c1 + 1):
                    if 0 <= nb < groups:
                        assert done[nb] is not None and it <= done[nb] <= it + 1, (groups, iterations, depth, st, nb, done[nb])
                for c in range(c0, c1 + 1):
                    done[c] = it + 1
            else:
                assert done[st[1]] == iterations
        assert done == [iterations] * groups
