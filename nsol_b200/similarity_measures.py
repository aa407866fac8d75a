"""Similarity measures of the reference API (``nsol.similarity_measures``), evaluated on the GPU.

SSD / MAE / MSE / RMSE / PSNR / NCC follow nsol/similarity_measures.py:26-120 and are computed
from one fused reduction pass (``nsol_similarity_stats``).  Called on numpy arrays they upload
both arrays and reduce on the device; called during a solver's probing pass (``_trace.Symbol``)
they describe themselves, which lets a solver evaluate them per iteration on the device-resident
iterate instead of copying every iterate to the host.

SSIM (``skimage.measure.compare_ssim``, removed from scikit-image), MI / NMI (histograms) and
Dice are evaluation-only utilities outside the hot path and are not implemented.
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


def _unavailable(name):
    def fn(*args, **kwargs):
        raise NotImplementedError("%s is an evaluation-only measure outside the CUDA hot path (SURVEY.md 2); "
                                  "it is not implemented in nsol_b200" % name)
    return fn


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

    structural_similarity = staticmethod(_unavailable("SSIM"))
    mutual_information = staticmethod(_unavailable("MI"))
    normalized_mutual_information = staticmethod(_unavailable("NMI"))
    dice_score = staticmethod(_unavailable("Dice"))

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
