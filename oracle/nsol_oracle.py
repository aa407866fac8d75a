"""CPU oracle for the NSoL proximal-solver hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy restatement of the reference algorithm
(gift-surg/NSoL v0.1.14, pure Python).  It is *not* part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / timed
CPU baseline.  ``nsol_b200`` never imports it and has no CPU fallback.

Parity pinning: the reference holds no golden vectors for this path (its tests
assert adjointness and x_scale invariance only).  The oracle is therefore
pinned by running the *unmodified reference itself* in the build container
(``oracle/gen_golden.py`` imports ``/root/reference/nsol`` with a 6-function
``pysitk`` stub) and committing its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those
fixtures (bit-exact for the primal-dual path, <=1e-12 for LSMR/ADMM).

Third-party arithmetic on the path, restated here from the published
algorithms:
  * ``scipy.ndimage.convolve`` (scipy 1.18.1; reference call sites
    nsol/linear_operators.py:65-66, 103-104, 198-199, 224-225, 244-245):
    direct weighted sum, ``mode="constant"`` pads with 0, ``mode="wrap"`` is
    periodic.
  * ``scipy.sparse.linalg.lsmr`` (scipy 1.18.1, Fong & Saunders 2011; reference
    call site nsol/tikhonov_linear_solver.py:149-154) and ``_sym_ortho``
    (S.-C. Choi's stable Givens rotation).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Layout conventions are the reference's: volumes are C-order
numpy arrays (z, y, x); "dx" acts on the last axis and is divided by
spacing[0], "dy" on axis -2 / spacing[1], "dz" on axis 0 / spacing[2]
(nsol/kernels.py:160-190, 240-286); the dual variable is SoA
``[Dx block | Dy block | Dz block]`` stacked on axis 0
(nsol/linear_operators.py:132, 140).
"""
from math import sqrt

import numpy as np

EPS = 1e-10  # nsol/definitions.py:11


# --------------------------------------------------------------------------
# Stencils: forward difference with zero boundary and its exact adjoint
# --------------------------------------------------------------------------
def _axis_of(component, ndim):
    """component 0 ("dx") -> last axis, 1 ("dy") -> axis -2, 2 ("dz") -> axis 0."""
    return ndim - 1 - component


def _shift_up(x, axis):
    """y[i] = x[i+1], y[last] = 0  (``mode="constant"``, cval=0)."""
    y = np.zeros_like(x)
    src = [slice(None)] * x.ndim
    dst = [slice(None)] * x.ndim
    src[axis] = slice(1, None)
    dst[axis] = slice(0, -1)
    y[tuple(dst)] = x[tuple(src)]
    return y


def _shift_down(x, axis):
    """y[i] = x[i-1], y[0] = 0."""
    y = np.zeros_like(x)
    src = [slice(None)] * x.ndim
    dst = [slice(None)] * x.ndim
    src[axis] = slice(0, -1)
    dst[axis] = slice(1, None)
    y[tuple(dst)] = x[tuple(src)]
    return y


def forward_difference(x, component, h=1.0):
    """D_k x  --  nsol/linear_operators.py:98-106 (dx), :193-201 (dy), :219-247 (dz)
    with masks nsol/kernels.py:102-107, 160-165, 172-177, 240-245, 253-258, 266-271.

    ``convolve(x, [1,-1]/h, mode="constant")`` evaluates
    ``fl(fl(w*x[i+1]) + fl((-w)*x[i]))`` with ``w = fl(1/h)`` and ``x[n] := 0``;
    bit-exact to scipy for any spacing (checked in tests/test_oracle_golden.py).
    """
    w = 1.0 / float(h)
    axis = _axis_of(component, x.ndim)
    return w * _shift_up(x, axis) + (-w) * x


def forward_difference_adj(y, component, h=1.0):
    """D_k^T y  --  same call sites, ``kernel_adj = -backward_difference``
    (nsol/kernels.py:109-112, 166-171, 184-190, 246-251, 259-264, 279-286):
    ``fl(fl(w*y[i-1]) + fl((-w)*y[i]))`` with ``y[-1] := 0``; no special last row.
    """
    w = 1.0 / float(h)
    axis = _axis_of(component, y.ndim)
    return w * _shift_down(y, axis) + (-w) * y


def _spacing(spacing, dim):
    s = np.atleast_1d(np.asarray(spacing, dtype=np.float64))
    if s.size == 1 and dim > 1:
        s = np.repeat(s, dim)
    if s.size != dim:
        # nsol/kernels.py:22-23
        raise ValueError("dimension of spacing and space must be the same")
    return s


def grad(x, spacing=None):
    """grad(x) = concatenate((Dx, Dy[, Dz])) on axis 0 -- nsol/linear_operators.py:121-144."""
    dim = x.ndim
    s = _spacing(np.ones(dim) if spacing is None else spacing, dim)
    return np.concatenate([forward_difference(x, k, s[k]) for k in range(dim)])


def grad_adj(p, spacing=None, dim=None):
    """grad_adj(p) = (Dx^T p_x + Dy^T p_y) + Dz^T p_z, blocks from
    ``np.array_split(p, dim)`` on axis 0 -- nsol/linear_operators.py:158-169."""
    dim = p.ndim if dim is None else dim
    s = _spacing(np.ones(dim) if spacing is None else spacing, dim)
    blocks = np.array_split(p, dim)
    out = forward_difference_adj(blocks[0], 0, s[0])
    for k in range(1, dim):
        out += forward_difference_adj(blocks[k], k, s[k])
    return out


# --------------------------------------------------------------------------
# Gaussian blur (periodic)
# --------------------------------------------------------------------------
def gaussian_kernel(dim, cov, spacing=None, alpha_cut=3):
    """Dense normalised Gaussian mask -- nsol/kernels.py:80-100 (1D), :120-158 (2D),
    :198-238 (3D).  Restates the reference's axis bookkeeping literally: the
    extent along numpy axis a is ``ceil(sqrt(cov[a,a])*alpha_cut/spacing[a])``
    while the quadratic form pairs axis a with ``(S cov^-1 S)`` row ``dim-1-a``.
    """
    s = _spacing(np.ones(dim) if spacing is None else spacing, dim)
    if dim == 1:
        cov = float(np.asarray(cov).reshape(-1)[0])
        x_max = np.ceil(np.sqrt(cov) * alpha_cut / s[0])
        points = np.arange(-x_max, x_max + 1, 1)
        cov_scale_inv = s[0] ** 2 / cov
        values = points * cov_scale_inv * points
        kernel = np.exp(-0.5 * values)
        return kernel / np.sum(kernel)
    cov = np.asarray(cov, dtype=np.float64)
    if cov.shape != (dim, dim):
        raise ValueError("Numpy array 'cov' must be of shape (%d,%d)" % (dim, dim))
    maxes = np.ceil(np.sqrt(cov.diagonal()) * alpha_cut / s)
    intervals = [np.arange(-m, m + 1, 1) for m in maxes]
    grids = np.meshgrid(*intervals, indexing="ij")
    # points = [Y, X] (2D) / [Z, Y, X] (3D): reversed grid order
    points = np.array([g.flatten() for g in grids[::-1]])
    S = np.diag(s)
    cov_scale_inv = S.dot(np.linalg.inv(cov)).dot(S)
    values = np.sum(points * cov_scale_inv.dot(points), 0)
    kernel = np.exp(-0.5 * values)
    kernel = kernel / np.sum(kernel)
    return kernel.reshape(*[iv.size for iv in intervals])


def convolve_wrap_dense(x, kernel):
    """``scipy.ndimage.convolve(x, kernel, mode="wrap")`` for an odd-sized mask --
    nsol/linear_operators.py:60-68.  out[i] = sum_k kernel[k] * x[(i - (k - c)) mod n]."""
    kernel = np.asarray(kernel, dtype=np.float64)
    out = np.zeros_like(x, dtype=np.float64)
    centre = [(n - 1) // 2 for n in kernel.shape]
    for idx in np.ndindex(*kernel.shape):
        wgt = kernel[idx]
        if wgt == 0.0:
            continue
        shifted = x
        for ax, (k, c) in enumerate(zip(idx, centre)):
            shifted = np.roll(shifted, k - c, axis=ax)
        out += wgt * shifted
    return out


def separable_taps(kernel, rtol=1e-13):
    """Rank-1 factorisation kernel == outer(t0, t1[, t2]) (diagonal covariance)
    or None if the mask is not separable.  Each factor is normalised to sum 1
    (the mask sums to 1, nsol/kernels.py:96, 151, 231)."""
    kernel = np.asarray(kernel, dtype=np.float64)
    taps = []
    for ax in range(kernel.ndim):
        other = tuple(a for a in range(kernel.ndim) if a != ax)
        t = kernel.sum(axis=other)
        taps.append(t / t.sum())
    rec = taps[0]
    for t in taps[1:]:
        rec = np.multiply.outer(rec, t)
    rec = rec * kernel.sum()
    if np.max(np.abs(rec - kernel)) > rtol * np.max(np.abs(kernel)):
        return None
    taps[0] = taps[0] * kernel.sum()
    return taps


def convolve_wrap_separable(x, taps):
    """Periodic separable convolution, one 1-D pass per axis (axis 0 first)."""
    out = np.asarray(x, dtype=np.float64)
    for ax, t in enumerate(taps):
        r = (t.size - 1) // 2
        acc = np.zeros_like(out)
        for k in range(t.size):
            acc += t[k] * np.roll(out, k - r, axis=ax)
        out = acc
    return out


# --------------------------------------------------------------------------
# Proximal maps -- nsol/proximal_operators.py
# --------------------------------------------------------------------------
def prox_ell1_denoising(x, tau, x0, x_scale=1.0):
    """nsol/proximal_operators.py:96-98."""
    x0 = x0 / float(x_scale)
    return x0 + np.maximum(np.abs(x - x0) - tau, 0) * np.sign(x - x0)


def prox_ell2_denoising(x, tau, x0, x_scale=1.0):
    """nsol/proximal_operators.py:118-120."""
    x0 = x0 / float(x_scale)
    return (x + tau * x0) / (1.0 + tau)


def prox_tv_conj(x, sigma):
    """nsol/proximal_operators.py:139-140 (element-wise / anisotropic projection)."""
    return x / np.maximum(1, np.abs(x))


def prox_huber_conj(x, sigma, gamma=0.05):
    """nsol/proximal_operators.py:157-159 (the reference divides its argument in
    place; callers pass a temporary)."""
    x = x / (1.0 + sigma * gamma)
    return x / np.maximum(1, np.abs(x))


def prox_tk1_conj(x, sigma):
    """BASELINE config 5's "TK1" as a primal-dual callable
    ``lambda q, s: q/(1+s)``: conjugate prox of 1/2 ||.||^2."""
    return x / (1.0 + sigma)


# --------------------------------------------------------------------------
# Primal-dual (Chambolle-Pock) -- nsol/primal_dual_solver.py
# --------------------------------------------------------------------------
def pd_initial_steps(alg_type, L2, lmbda, huber_alpha=0.05):
    """(tau0, sigma0, gamma) -- nsol/primal_dual_solver.py:278-283 (ALG2),
    :321-337 (ALG3; theta returned in the gamma slot), :374-379 (AHMOD)."""
    if alg_type == "ALG2":
        tau0 = 1.0 / np.sqrt(L2)
        sigma0 = 1.0 / (L2 * tau0)
        return tau0, sigma0, 0.35 * lmbda
    if alg_type == "ALG3":
        gamma = lmbda
        delta = huber_alpha
        mu = 2.0 * np.sqrt(gamma * delta / L2)
        theta = 1.0 / (1.0 + mu)
        sigma = mu / (2.0 * delta)
        tau = mu / (2.0 * gamma)
        return tau, sigma, theta
    if alg_type == "ALG2_AHMOD":
        tau0 = 0.02
        sigma0 = 4.0 / (L2 * tau0)
        return tau0, sigma0, 0.35 * lmbda
    raise KeyError(alg_type)


def pd_update_steps(alg_type, L2, gamma, tau_n, sigma_n):
    """(theta, tau, sigma) -- nsol/primal_dual_solver.py:302-306 (ALG2),
    :356-358 (ALG3), :398-403 (AHMOD, theta = 0)."""
    if alg_type == "ALG3":
        return gamma, tau_n, sigma_n
    theta_n = 1.0 / np.sqrt(1.0 + 2.0 * gamma * tau_n)
    tau_n = tau_n * theta_n
    sigma_n = sigma_n / theta_n
    if alg_type == "ALG2_AHMOD":
        return 0.0, tau_n, sigma_n
    return theta_n, tau_n, sigma_n


def pd_schedule(alg_type, L2, alpha, iterations):
    """Per-iteration scalars (sigma_n, tau_n, tau_n*lmbda, theta_n) exactly as the
    loop nsol/primal_dual_solver.py:222-253 produces them (float64 host math)."""
    lmbda = 1.0 / alpha
    tau_n, sigma_n, gamma = pd_initial_steps(alg_type, float(L2), lmbda)
    rows = np.zeros((iterations, 4))
    for i in range(iterations):
        sig_i, tau_i = sigma_n, tau_n
        theta_n, tau_n, sigma_n = pd_update_steps(alg_type, float(L2), gamma, tau_n, sigma_n)
        rows[i] = (sig_i, tau_i, tau_i * lmbda, theta_n)
    return rows


_PROX_G = {"TV": prox_tv_conj, "HUBER": prox_huber_conj, "TK1": prox_tk1_conj}
_PROX_F = {"L1": prox_ell1_denoising, "L2": prox_ell2_denoising}


def primal_dual_denoise(b, shape, reg="TV", data="L2", alpha=0.01, L2=8.0, iterations=10,
                        x_scale=1.0, alg_type="ALG2", spacing=None, x0=None,
                        huber_gamma=0.05, keep_iterates=False):
    """Chambolle-Pock loop nsol/primal_dual_solver.py:215-263 with the denoising
    wiring of nsol/application/run_denoising.py:95-154:
    B = grad, B_conj = grad_adj, prox_f in {ell1, ell2}(.., x0=b, x_scale),
    prox_g_conj in {tv, huber}.  ``b``/``x0`` are unscaled 1-D observations
    (nsol/solver.py:35-41 divides by x_scale; :117-118 multiplies back).
    Returns the 1-D result of ``get_x()`` (and every iterate if asked, as the
    Observer would see them, nsol/primal_dual_solver.py:218-219, 260-261).
    """
    shape = tuple(int(s) for s in shape)
    dim = len(shape)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    x0 = b if x0 is None else np.asarray(x0, dtype=np.float64).reshape(-1)
    x_scale = float(x_scale)
    zshape = (dim * shape[0],) + shape[1:]
    B = lambda x: grad(x.reshape(shape), spacing).reshape(-1)
    B_conj = lambda p: grad_adj(p.reshape(zshape), spacing, dim).reshape(-1)
    if reg == "HUBER":
        prox_g = lambda q, s: prox_huber_conj(q, s, huber_gamma)
    else:
        prox_g = _PROX_G[reg]
    prox_f = lambda x, t: _PROX_F[data](x, t, x0=b, x_scale=x_scale)

    lmbda = 1.0 / float(alpha)
    tau_n, sigma_n, gamma = pd_initial_steps(alg_type, float(L2), lmbda)
    x_n = np.array(x0) / x_scale
    x_mean = np.array(x_n)
    p_n = 0
    iterates = [x_n * x_scale] if keep_iterates else None
    for _ in range(iterations):
        p_n = prox_g(p_n + sigma_n * B(x_mean), sigma_n)
        x_np1 = prox_f(x_n - tau_n * B_conj(p_n), tau_n * lmbda)
        theta_n, tau_n, sigma_n = pd_update_steps(alg_type, float(L2), gamma, tau_n, sigma_n)
        x_mean = x_np1 + theta_n * (x_np1 - x_n)
        x_n = x_np1
        if keep_iterates:
            iterates.append(x_n * x_scale)
    if keep_iterates:
        return x_n * x_scale, iterates
    return x_n * x_scale


def primal_dual_denoise_ndimage(b, shape, reg="TV", data="L2", alpha=0.01, L2=8.0,
                                iterations=10, x_scale=1.0, alg_type="ALG2"):
    """Same loop, but with the operator closures built the way the reference
    builds them (``scipy.ndimage.convolve`` + ``np.concatenate`` /
    ``np.array_split``; nsol/linear_operators.py:98-169, unit spacing).  This is
    the form ``bench.py`` times as the CPU baseline: it has the reference's own
    per-iteration cost profile (2d ndimage passes + ~15 numpy temporaries).
    """
    import scipy.ndimage
    shape = tuple(int(s) for s in shape)
    dim = len(shape)
    fwd, adj = [], []
    for k in range(dim):
        kshape = [1] * dim
        kshape[dim - 1 - k] = 2
        kf = np.array([1.0, -1.0]).reshape(kshape)
        kshape[dim - 1 - k] = 3
        ka = -np.array([0.0, 1.0, -1.0]).reshape(kshape)
        fwd.append(lambda x, kf=kf: scipy.ndimage.convolve(x, kf, mode="constant"))
        adj.append(lambda x, ka=ka: scipy.ndimage.convolve(x, ka, mode="constant"))
    zshape = (dim * shape[0],) + shape[1:]

    def B(x):
        x = x.reshape(shape)
        if dim == 1:
            return fwd[0](x).flatten()
        return np.concatenate([f(x) for f in fwd]).flatten()

    def B_conj(p):
        parts = np.array_split(p.reshape(zshape), dim)
        out = adj[0](parts[0])
        for k in range(1, dim):
            out += adj[k](parts[k])
        return out.flatten()

    b = np.asarray(b, dtype=np.float64).reshape(-1)
    x_scale = float(x_scale)
    prox_g = _PROX_G[reg]
    prox_f = lambda x, t: _PROX_F[data](x, t, x0=b, x_scale=x_scale)
    lmbda = 1.0 / float(alpha)
    tau_n, sigma_n, gamma = pd_initial_steps(alg_type, float(L2), lmbda)
    x_n = np.array(b) / x_scale
    x_mean = np.array(x_n)
    p_n = 0
    for _ in range(iterations):
        p_n = prox_g(p_n + sigma_n * B(x_mean), sigma_n)
        x_np1 = prox_f(x_n - tau_n * B_conj(p_n), tau_n * lmbda)
        theta_n, tau_n, sigma_n = pd_update_steps(alg_type, float(L2), gamma, tau_n, sigma_n)
        x_mean = x_np1 + theta_n * (x_np1 - x_n)
        x_n = x_np1
    return x_n * x_scale


# --------------------------------------------------------------------------
# LSMR -- scipy/sparse/linalg/_isolve/lsmr.py:29-494 (scipy 1.18.1), damp=0
# --------------------------------------------------------------------------
def sym_ortho(a, b):
    """Stable Givens rotation -- scipy/sparse/linalg/_isolve/lsqr.py:62-94."""
    if b == 0:
        return np.sign(a), 0, abs(a)
    elif a == 0:
        return 0, np.sign(b), abs(b)
    elif abs(b) > abs(a):
        tau = a / b
        s = np.sign(b) / sqrt(1 + tau * tau)
        c = s * tau
        r = b / s
    else:
        tau = b / a
        c = np.sign(a) / sqrt(1 + tau * tau)
        s = c * tau
        r = a / c
    return c, s, r


def lsmr(matvec, rmatvec, b, n, maxiter, atol=0.0, btol=0.0, conlim=1e8, damp=0.0, norm=None):
    """LSMR (Fong & Saunders) as scipy runs it for the reference call
    ``lsmr(A, b, maxiter=iter_max, atol=0, btol=0)`` (cold start, x0=None),
    nsol/tikhonov_linear_solver.py:149-154.  Returns (x, istop, itn).
    ``norm``: 2-norm callable (default np.linalg.norm); the z-slab tests pass an all-reduced norm so that
    the same recurrences run on the local slab of every rank."""
    norm = norm or np.linalg.norm
    b = np.asarray(b, dtype=np.float64)
    u = b
    normb = norm(b)
    x = np.zeros(n)
    beta = normb.copy()
    if beta > 0:
        u = (1 / beta) * u
        v = rmatvec(u)
        alpha = norm(v)
    else:
        v = np.zeros(n)
        alpha = 0
    if alpha > 0:
        v = (1 / alpha) * v

    itn = 0
    zetabar = alpha * beta
    alphabar = alpha
    rho = 1
    rhobar = 1
    cbar = 1
    sbar = 0
    h = v.copy()
    hbar = np.zeros(n)

    betadd = beta
    betad = 0
    rhodold = 1
    tautildeold = 0
    thetatilde = 0
    zeta = 0
    d = 0

    normA2 = alpha * alpha
    maxrbar = 0
    minrbar = 1e100
    istop = 0
    ctol = 0
    if conlim > 0:
        ctol = 1 / conlim
    normar = alpha * beta
    if normar == 0:
        return x, istop, itn
    if normb == 0:
        x[()] = 0
        return x, istop, itn

    while itn < maxiter:
        itn += 1
        u *= -alpha
        u += matvec(v)
        beta = norm(u)
        if beta > 0:
            u *= (1 / beta)
            v *= -beta
            v += rmatvec(u)
            alpha = norm(v)
            if alpha > 0:
                v *= (1 / alpha)

        chat, shat, alphahat = sym_ortho(alphabar, damp)
        rhoold = rho
        c, s, rho = sym_ortho(alphahat, beta)
        thetanew = s * alpha
        alphabar = c * alpha

        rhobarold = rhobar
        zetaold = zeta
        thetabar = sbar * rho
        rhotemp = cbar * rho
        cbar, sbar, rhobar = sym_ortho(cbar * rho, thetanew)
        zeta = cbar * zetabar
        zetabar = -sbar * zetabar

        hbar *= -(thetabar * rho / (rhoold * rhobarold))
        hbar += h
        x += (zeta / (rho * rhobar)) * hbar
        h *= -(thetanew / rho)
        h += v

        betaacute = chat * betadd
        betacheck = -shat * betadd
        betahat = c * betaacute
        betadd = -s * betaacute

        thetatildeold = thetatilde
        ctildeold, stildeold, rhotildeold = sym_ortho(rhodold, thetabar)
        thetatilde = stildeold * rhobar
        rhodold = ctildeold * rhobar
        betad = -stildeold * betad + ctildeold * betahat

        tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold
        taud = (zeta - thetatilde * tautildeold) / rhodold
        d = d + betacheck * betacheck
        normr = sqrt(d + (betad - taud) ** 2 + betadd * betadd)

        normA2 = normA2 + beta * beta
        normA = sqrt(normA2)
        normA2 = normA2 + alpha * alpha

        maxrbar = max(maxrbar, rhobarold)
        if itn > 1:
            minrbar = min(minrbar, rhobarold)
        condA = max(maxrbar, rhotemp) / min(minrbar, rhotemp)

        normar = abs(zetabar)
        normx = norm(x)

        test1 = normr / normb
        if (normA * normr) != 0:
            test2 = normar / (normA * normr)
        else:
            test2 = np.inf
        test3 = 1 / condA
        t1 = test1 / (1 + normA * normx / normb)
        rtol = btol + atol * normA * normx / normb

        if itn >= maxiter:
            istop = 7
        if 1 + test3 <= 1:
            istop = 6
        if 1 + test2 <= 1:
            istop = 5
        if 1 + t1 <= 1:
            istop = 4
        if test3 <= ctol:
            istop = 3
        if test2 <= atol:
            istop = 2
        if test1 <= rtol:
            istop = 1
        if istop > 0:
            break
    return x, istop, itn


# --------------------------------------------------------------------------
# Tikhonov (lsmr / linear branch) -- nsol/tikhonov_linear_solver.py
# --------------------------------------------------------------------------
def tikhonov_lsmr(A, A_adj, B, B_adj, b, x0, alpha=0.01, b_reg=0, iter_max=10,
                  x_scale=1.0, bounds=(0, np.inf), norm=None):
    """``TikhonovLinearSolver.run()`` with minimizer="lsmr", data_loss="linear":
    augmented system [A; sqrt(alpha) B] x = [b; sqrt(alpha) b_reg]
    (nsol/tikhonov_linear_solver.py:226-274), LSMR cold start (:149-154), clip to
    bounds (:156-158); scaling nsol/linear_solver.py:80-89, nsol/solver.py:35-41,
    :117-118.  All vectors 1-D.  Returns ``get_x()``."""
    x_scale = float(x_scale)
    x0s = np.asarray(x0, dtype=np.float64) / x_scale
    bs = np.asarray(b, dtype=np.float64) / x_scale
    b_regs = b_reg / x_scale
    n = x0s.size
    if alpha > EPS:
        sa = np.sqrt(alpha)
        m_up = bs.size
        fw = lambda x: np.concatenate((A(x), sa * B(x)))
        bw = lambda y: A_adj(y[:m_up]) + sa * B_adj(y[m_up:])
        rhs = np.zeros(fw(x0s).size)
        rhs[0:m_up] = bs
        rhs[m_up:] = sa * b_regs
    else:
        fw, bw, rhs = A, A_adj, bs
    x = lsmr(fw, bw, np.array(rhs), n, maxiter=iter_max, norm=norm)[0]
    if bounds is not None:
        x = np.clip(x, bounds[0], bounds[1])
    return x * x_scale


# --------------------------------------------------------------------------
# ADMM TV-L2 -- nsol/admm_linear_solver.py
# --------------------------------------------------------------------------
def admm_shrink_iso(t, ell, dim):
    """``ADMMLinearSolver._prox_g`` -- nsol/admm_linear_solver.py:239-253 (+ :268-309):
    isotropic soft threshold of the d SoA blocks of t."""
    parts = np.array_split(t, dim)
    tmp = parts[0] ** 2
    for k in range(1, dim):
        tmp += parts[k] ** 2
    t_norm = np.sqrt(tmp)
    ind = t_norm > ell
    v = np.zeros_like(t)
    m = parts[0].shape[0]
    for k in range(dim):
        vk = v[k * m:(k + 1) * m]
        soft = np.maximum(np.abs(t_norm[ind]) - ell, 0) * np.sign(t_norm[ind])
        vk[ind] = soft * parts[k][ind] / t_norm[ind]
    return v


def admm_tv(A, A_adj, B, B_adj, b, x0, dim, alpha=0.01, rho=0.5, iterations=10,
            iter_max=10, x_scale=1.0, keep_iterates=False, norm=None, b_reg=0):
    """``ADMMLinearSolver.run()`` -- nsol/admm_linear_solver.py:165-237:
    v = B(x0) - b_reg, w = 0; per iteration x <- Tikhonov/LSMR solve with
    b_reg' = v - w + b_reg and weight rho (x_scale=1, bounds (0, inf)), t = B(x) + w - b_reg,
    v = shrink_iso(t, alpha/rho), w = t - v.  (b_reg is divided by x_scale, :100.)"""
    x_scale = float(x_scale)
    x = np.asarray(x0, dtype=np.float64) / x_scale
    bs = np.asarray(b, dtype=np.float64) / x_scale
    c = b_reg / x_scale
    v = B(x) - c
    w = np.zeros_like(v)
    iterates = [x * x_scale] if keep_iterates else None
    for _ in range(iterations):
        x = tikhonov_lsmr(A, A_adj, B, B_adj, bs, x, alpha=rho, b_reg=v - w + c,
                          iter_max=iter_max, x_scale=1.0, norm=norm)
        t = B(x) + w - c
        v = admm_shrink_iso(t, alpha / rho, dim)
        w = t - v
        if keep_iterates:
            iterates.append(x * x_scale)
    if keep_iterates:
        return x * x_scale, iterates
    return x * x_scale


def primal_dual_deconvolve(A, A_adj, D, D_adj, b, x0, dim, reg="TV", alpha=0.01, L2=8.0, iterations=10,
                           iter_max=10, x_scale=1.0, alg_type="ALG2", huber_gamma=0.05):
    """PrimalDualSolver with prox_f = prox_linear_least_squares -- the default TV-L2 / Huber-L2
    deconvolution wiring, nsol/deconvolution_solver_parameter_study_interface.py:255-280, 303-325 and
    nsol/proximal_operators.py:44-78 (the prox divides b and x0 by x_scale and hands them to a
    Tikhonov solver that divides by x_scale again; restated literally)."""
    x_scale = float(x_scale)
    b = np.asarray(b, dtype=np.float64)
    ident = lambda v: v.reshape(-1)

    def prox_f(x, tau):
        return tikhonov_lsmr(A, A_adj, ident, ident, b / x_scale, np.asarray(x0) / x_scale, alpha=1.0 / tau, b_reg=x,
                             iter_max=iter_max, x_scale=x_scale)
    prox_g = (lambda q, s: prox_huber_conj(q, s, huber_gamma)) if reg == "HUBER" else _PROX_G[reg]
    lmbda = 1.0 / float(alpha)
    tau_n, sigma_n, gamma = pd_initial_steps(alg_type, float(L2), lmbda)
    x_n = np.asarray(x0, dtype=np.float64) / x_scale
    x_mean = np.array(x_n)
    p_n = 0
    for _ in range(iterations):
        p_n = prox_g(p_n + sigma_n * D(x_mean), sigma_n)
        x_np1 = prox_f(x_n - tau_n * D_adj(p_n), tau_n * lmbda)
        theta_n, tau_n, sigma_n = pd_update_steps(alg_type, float(L2), gamma, tau_n, sigma_n)
        x_mean = x_np1 + theta_n * (x_np1 - x_n)
        x_n = x_np1
    return x_n * x_scale


def deconvolution_operators(shape, cov, spacing=None, alpha_cut=3, separable=True):
    """1-D wrapped A, A_adj, D, D_adj as nsol/application/run_deconvolution.py:109-129
    builds them (A_adj reuses the same mask, nsol/linear_operators.py:63)."""
    shape = tuple(int(s) for s in shape)
    dim = len(shape)
    kernel = gaussian_kernel(dim, cov, spacing, alpha_cut)
    taps = separable_taps(kernel) if separable else None
    if taps is not None:
        blur = lambda x: convolve_wrap_separable(x, taps)
    else:
        blur = lambda x: convolve_wrap_dense(x, kernel)
    zshape = (dim * shape[0],) + shape[1:]
    A = lambda x: blur(x.reshape(shape)).reshape(-1)
    D = lambda x: grad(x.reshape(shape), spacing).reshape(-1)
    D_adj = lambda p: grad_adj(p.reshape(zshape), spacing, dim).reshape(-1)
    return A, A, D, D_adj


# --------------------------------------------------------------------------
# Noise -- nsol/noise.py (legacy numpy RNG; used to build benchmark inputs)
# --------------------------------------------------------------------------
def add_gaussian_noise(data, noise_level=0.01, seed=1):
    """nsol/noise.py:28, 51-55."""
    np.random.seed(seed=seed)
    data = np.array(data, dtype=np.float64)
    data += noise_level * data.max() * np.random.normal(size=data.shape, loc=0, scale=1)
    return data


def add_salt_and_pepper_noise(data, salt_vs_pepper=0.5, amount=0.1, seed=1):
    """nsol/noise.py:28, 86-109."""
    np.random.seed(seed=seed)
    data = np.array(data, dtype=np.float64)
    val_salt, val_pepper = data.max(), data.min()
    shape = data.shape
    data = data.flatten()
    size = int(amount * data.size)
    samples_all = np.random.choice(np.arange(0, data.size), size=size, replace=False)
    n_white = int(salt_vs_pepper * samples_all.size)
    data[samples_all[0:n_white]] = val_salt
    data[samples_all[n_white:]] = val_pepper
    return data.reshape(*shape)


# --------------------------------------------------------------------------
# Similarity measures used by the 3-decimal parity check (north_star)
# --------------------------------------------------------------------------
def psnr(x, x_ref):
    """nsol/similarity_measures.py:99-101 (with MSE :56-58)."""
    mse = np.sum(np.square(x - x_ref)) / float(x.size)
    return 10 * np.log10(np.max(x_ref) ** 2 / mse)


def ncc(x, x_ref):
    """nsol/similarity_measures.py:113-120 (ddof=1 std, divided by x.size)."""
    a = x - np.mean(x)
    r = x_ref - np.mean(x_ref)
    return np.sum(a * r) / (float(x.size) * np.std(x, ddof=1) * np.std(x_ref, ddof=1))


def ssim_1d(x, x_ref, win=7, K1=0.01, K2=0.03):
    """Restatement of ``skimage.measure.compare_ssim(x, x_ref)`` with defaults on
    the *flattened* arrays (nsol/similarity_measures.py:135-136): Wang 2004,
    uniform 7-sample window, sample covariance, data_range from dtype float ->
    skimage used (max - min of the dtype range) = 2 for floats; parity here only
    needs the *same* restatement applied to both outputs (unpinned)."""
    import scipy.ndimage
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    y = np.asarray(x_ref, dtype=np.float64).reshape(-1)
    data_range = float(y.max() - y.min())
    filt = lambda a: scipy.ndimage.uniform_filter1d(a, win, mode="reflect")
    cov_norm = win / (win - 1.0)
    ux, uy = filt(x), filt(y)
    uxx, uyy, uxy = filt(x * x), filt(y * y), filt(x * y)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (K1 * data_range) ** 2
    C2 = (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win - 1) // 2
    return float(S[pad:-pad].mean())
