"""Similarity measures of the reference API (``nsol.similarity_measures``), evaluated on the GPU.

SSD / MAE / MSE / RMSE / PSNR / NCC follow nsol/similarity_measures.py:26-120 and are computed
from one fused reduction pass (``nsol_similarity_stats``).  Called on numpy arrays they upload
both arrays and reduce on the device; called during a solver's probing pass (``_trace.Symbol``)
they describe themselves, which lets a solver evaluate them per iteration on the device-resident
iterate instead of copying every iterate to the host.

SSIM, Shannon / joint entropy, MI / NMI and Dice (nsol/similarity_measures.py:135-264) are
evaluation-only utilities outside the hot path: they are evaluated on the HOST
with numpy, after the solve, on the iterates the Observer holds.  SSIM restates
``skimage.measure.compare_ssim(x, x_ref)`` with its defaults (Wang et al. 2004: uniform
7-sample window, sample covariance, K1 = 0.01, K2 = 0.03) because that function no longer
exists in scikit-image; its ``data_range`` is taken from the reference image (SURVEY.md 8c:
SSIM parity is unpinned).
"""
import ctypes as C

import numpy as np

from nsol_b200 import _lib
from nsol_b200._trace import is_symbol

DEVICE_MEASURES = ("SSD", "MAE", "MSE", "RMSE", "PSNR", "NCC")


class MeasureRequest(object):
    """What a measure callable returns when probed with a Symbol."""

    def __init__(self, kind, x_ref):
        self.kind = kind
        self.x_ref = x_ref


def from_stats(kind, st, n):
    """Measure value from the 8 sums (sum y, y^2, r, r^2, y r, (y-r)^2, |y-r|, max r)."""
    sy, syy, sr, srr, syr, ssd, sad, rmax = [float(v) for v in st]
    n = float(n)
    if kind == "SSD":
        return ssd
    if kind == "MAE":
        return sad / n
    if kind == "MSE":
        return ssd / n
    if kind == "RMSE":
        return np.sqrt(ssd / n)
    if kind == "PSNR":
        with np.errstate(divide="ignore"):
            return float(10 * np.log10(np.float64(rmax ** 2) / np.float64(ssd / n)))
    if kind == "NCC":
        cov = syr - sy * sr / n
        var_y = (syy - sy * sy / n) / (n - 1.0)
        var_r = (srr - sr * sr / n) / (n - 1.0)
        return cov / (n * np.sqrt(var_y) * np.sqrt(var_r))
    raise KeyError(kind)


def device_stats(ctx, dtype_code, n, x_dev_ptr, scale, xref_buf):
    out = np.empty(8, dtype=np.float64)
    ctx.check(ctx.lib.nsol_similarity_stats(ctx.handle, dtype_code, n, x_dev_ptr, float(scale), xref_buf.ptr,
                                            out.ctypes.data, None))
    return out


def _measure(kind, x, x_ref):
    if is_symbol(x):
        return MeasureRequest(kind, x_ref)
    x = np.ascontiguousarray(x, dtype=np.float64)
    x_ref = np.ascontiguousarray(x_ref, dtype=np.float64)
    if x.shape != x_ref.shape:
        raise ValueError("Input data shapes do not match")
    ctx = _lib.context()
    dx = ctx.device_alloc(max(x.nbytes, 8)).upload(x)
    dr = ctx.device_alloc(max(x_ref.nbytes, 8)).upload(x_ref)
    try:
        st = device_stats(ctx, _lib.F64, x.size, dx.ptr, 1.0, dr)
    finally:
        dx.free()
        dr.free()
    return from_stats(kind, st, x.size)


def _host(x):
    if is_symbol(x):
        raise TypeError("this measure is evaluated on the host, on stored iterates (use a storing Observer)")
    return np.asarray(x, dtype=np.float64)


def _window_mean(a, win):
    """Mean over a centred window of ``win`` samples at every position where the window fits."""
    c = np.concatenate(([0.0], np.cumsum(a)))
    return (c[win:] - c[:-win]) / float(win)


def _entropy(hist):
    prob = hist[hist > 0] / float(np.sum(hist))
    return float(-np.sum(prob * np.log(prob)))


class SimilarityMeasures(object):

    @staticmethod
    def sum_of_absolute_differences(x, x_ref):
        return _measure("MAE", x, x_ref) * float(np.size(x_ref))

    @staticmethod
    def mean_absolute_error(x, x_ref):
        return _measure("MAE", x, x_ref)

    @staticmethod
    def sum_of_squared_differences(x, x_ref):
        return _measure("SSD", x, x_ref)

    @staticmethod
    def mean_squared_error(x, x_ref):
        return _measure("MSE", x, x_ref)

    @staticmethod
    def root_mean_square_error(x, x_ref):
        return _measure("RMSE", x, x_ref)

    @staticmethod
    def peak_signal_to_noise_ratio(x, x_ref):
        return _measure("PSNR", x, x_ref)

    @staticmethod
    def normalized_cross_correlation(x, x_ref):
        return _measure("NCC", x, x_ref)

    @staticmethod
    def structural_similarity(x, x_ref, win=7, K1=0.01, K2=0.03):
        """Mean structural similarity of the flattened arrays (nsol/similarity_measures.py:135-136), host."""
        a, r = _host(x).reshape(-1), _host(x_ref).reshape(-1)
        if a.shape != r.shape:
            raise ValueError("Input data shapes do not match")
        if a.size < win:
            raise ValueError("structural_similarity needs at least %d samples" % win)
        data_range = float(r.max() - r.min())
        norm = win / (win - 1.0)
        ma, mr = _window_mean(a, win), _window_mean(r, win)
        va = norm * (_window_mean(a * a, win) - ma * ma)
        vr = norm * (_window_mean(r * r, win) - mr * mr)
        var = norm * (_window_mean(a * r, win) - ma * mr)
        c1, c2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
        ssim = ((2 * ma * mr + c1) * (2 * var + c2)) / ((ma * ma + mr * mr + c1) * (va + vr + c2))
        return float(np.mean(ssim))

    @staticmethod
    def shannon_entropy(x, bins=100):
        """H(X) = -sum p ln p over a histogram of the flattened array (nsol/similarity_measures.py:154-165), host."""
        return _entropy(np.histogram(_host(x), bins=bins)[0])

    @staticmethod
    def joint_entropy(x, x_ref, bins=100):
        """nsol/similarity_measures.py:183-191, host."""
        return _entropy(np.histogram2d(_host(x).reshape(-1), _host(x_ref).reshape(-1), bins=bins)[0])

    @staticmethod
    def mutual_information(x, x_ref, bins=100):
        """H(X) + H(Y) - H(X,Y)  (nsol/similarity_measures.py:212-216), host."""
        sm = SimilarityMeasures
        return sm.shannon_entropy(x, bins) + sm.shannon_entropy(x_ref, bins) - sm.joint_entropy(x, x_ref, bins)

    @staticmethod
    def normalized_mutual_information(x, x_ref, bins=100):
        """(H(X) + H(Y)) / H(X,Y)  (nsol/similarity_measures.py:235-239), host."""
        sm = SimilarityMeasures
        return (sm.shannon_entropy(x, bins) + sm.shannon_entropy(x_ref, bins)) / sm.joint_entropy(x, x_ref, bins)

    @staticmethod
    def dice_score(x, x_ref):
        """2 |A and B| / (|A| + |B|) of two boolean arrays (nsol/similarity_measures.py:255-264), host."""
        x, x_ref = np.asarray(x), np.asarray(x_ref)
        if x.dtype != np.bool_ or x_ref.dtype != np.bool_:
            raise ValueError("x and x_ref need to be of type boolean")
        return 2.0 * float(np.sum(x & x_ref)) / float(np.sum(x) + np.sum(x_ref))

    # nsol/similarity_measures.py:267-277
    similarity_measures = {
        "SSD": sum_of_squared_differences.__func__,
        "MAE": mean_absolute_error.__func__,
        "MSE": mean_squared_error.__func__,
        "RMSE": root_mean_square_error.__func__,
        "PSNR": peak_signal_to_noise_ratio.__func__,
        "NCC": normalized_cross_correlation.__func__,
        "SSIM": structural_similarity.__func__,
        "MI": mutual_information.__func__,
        "NMI": normalized_mutual_information.__func__,
    }
    UNDEF = {k: np.nan for k in ("SSD", "MAE", "MSE", "RMSE", "PSNR", "SSIM", "NCC", "MI", "NMI")}
