"""Convolution masks of the reference API (``nsol.kernels``): finite-difference
stencils and normalised Gaussian masks as small host-side numpy arrays.

Mirrors ``Kernels1D/2D/3D`` of the reference (nsol/kernels.py:15-286): same class
and method names, same axis conventions ("dx" = last numpy axis / spacing[0],
"dy" = axis -2 / spacing[1], "dz" = axis 0 / spacing[2]) and the same
``ValueError``s.  These arrays are *parameters* (a few taps); the convolutions
themselves run on the GPU (linear_operators.py).
"""
import numpy as np


class Kernels(object):
    """Dimension-generic implementation; Kernels1D/2D/3D fix ``dimension``."""

    def __init__(self, dimension, spacing):
        spacing = np.atleast_1d(spacing).astype(float)
        if spacing.size != dimension:
            # nsol/kernels.py:22-23
            raise ValueError("dimension of spacing and space must be the same")
        self._dimension = dimension
        self._spacing = spacing

    def get_dimension(self):
        return self._dimension

    def get_spacing(self):
        return self._spacing

    # -- Gaussian ---------------------------------------------------------
    def get_gaussian(self, cov, alpha_cut=3):
        """Normalised Gaussian mask truncated at ``alpha_cut`` standard deviations
        (nsol/kernels.py:80-100, 120-158, 198-238).  The reference pairs numpy
        axis a with row dim-1-a of ``S cov^-1 S`` while taking the extent of axis
        a from ``cov[a, a]`` and ``spacing[a]``; that bookkeeping is kept."""
        d = self._dimension
        if d == 1:
            var = float(np.asarray(cov, dtype=float).reshape(-1)[0])
            half = np.ceil(np.sqrt(var) * alpha_cut / self._spacing[0])
            t = np.arange(-half, half + 1, 1)
            mask = np.exp(-0.5 * (t * (self._spacing[0] ** 2 / var) * t))
            return mask / np.sum(mask)
        cov = np.asarray(cov)
        if cov.shape != (d, d):
            raise ValueError("Numpy array 'cov' must be of shape (%d,%d)" % (d, d))
        half = np.ceil(np.sqrt(cov.diagonal()) * alpha_cut / self._spacing)
        axes = [np.arange(-h, h + 1, 1) for h in half]
        mesh = np.meshgrid(*axes, indexing="ij")
        pts = np.array([m.flatten() for m in reversed(mesh)])
        S = np.diag(self._spacing)
        Q = S.dot(np.linalg.inv(cov)).dot(S)
        mask = np.exp(-0.5 * np.sum(pts * Q.dot(pts), 0))
        mask = mask / np.sum(mask)
        return mask.reshape([a.size for a in axes])

    # -- finite differences -------------------------------------------------
    def _difference(self, component, taps):
        """taps laid along the numpy axis of derivative ``component``, divided by its spacing."""
        d = self._dimension
        if component >= d:
            raise AttributeError("d%s not defined in %dD" % ("xyz"[component], d))
        shape = [1] * d
        shape[d - 1 - component] = len(taps)
        mask = np.asarray(taps, dtype=float).reshape(shape)
        if d == 1:
            mask = np.asarray(taps)
        return mask / self._spacing[component]

    def get_dx_forward_difference(self):
        return self._difference(0, [1, -1])

    def get_dx_backward_difference(self):
        return self._difference(0, [0, 1, -1])

    def get_dy_forward_difference(self):
        return self._difference(1, [1, -1])

    def get_dy_backward_difference(self):
        return self._difference(1, [0, 1, -1])

    def get_dz_forward_difference(self):
        return self._difference(2, [1, -1])

    def get_dz_backward_difference(self):
        return self._difference(2, [0, 1, -1])


class Kernels1D(Kernels):
    def __init__(self, spacing=1):
        Kernels.__init__(self, dimension=1, spacing=spacing)


class Kernels2D(Kernels):
    def __init__(self, spacing=np.ones(2)):
        Kernels.__init__(self, dimension=2, spacing=spacing)


class Kernels3D(Kernels):
    def __init__(self, spacing=np.ones(3)):
        Kernels.__init__(self, dimension=3, spacing=spacing)


def separable_taps(mask, rtol=1e-13):
    """Rank-1 factors of a mask (one tap vector per numpy axis) or None.

    For a diagonal covariance the reference's dense mask is exactly
    ``outer(g0, g1[, g2])`` (SURVEY.md headline fact 5), which lets the GPU run
    d one-dimensional passes instead of a (2r+1)^d-tap dense convolution."""
    mask = np.asarray(mask, dtype=np.float64)
    if mask.ndim == 1:
        return [mask]
    total = mask.sum()
    if total == 0:
        return None
    taps = []
    for ax in range(mask.ndim):
        t = mask.sum(axis=tuple(a for a in range(mask.ndim) if a != ax))
        taps.append(t / total)
    rec = taps[0]
    for t in taps[1:]:
        rec = np.multiply.outer(rec, t)
    if np.max(np.abs(rec * total - mask)) > rtol * np.max(np.abs(mask)):
        return None
    taps[0] = taps[0] * total
    return taps
