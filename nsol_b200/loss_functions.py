"""Data-loss functions rho(f^2) of the reference API (``nsol.loss_functions.LossFunctions``,
nsol/loss_functions.py:25-266) -- host-side numpy, diagnostics only.

Only the ``linear`` loss reaches the GPU path (the lsmr branch rejects every other loss,
nsol/tikhonov_linear_solver.py:122-124); the robust losses exist here so that the "Data" cost
measure of the deconvolution interface can be evaluated for any ``data_loss`` string, with the
definitions scipy.optimize.least_squares documents (rho_C(f2) = C^2 rho(f2 / C^2))."""
import numpy as np


def _scaled(rho):
    def loss(f2, f_scale=1.):
        c2 = float(f_scale) ** 2
        return c2 * rho(np.asarray(f2, dtype=np.float64) / c2)
    return loss


class LossFunctions(object):

    linear = staticmethod(lambda f2, f_scale=1.: np.asarray(f2, dtype=np.float64))
    soft_l1 = staticmethod(_scaled(lambda z: 2. * (np.sqrt(1. + z) - 1.)))
    huber = staticmethod(_scaled(lambda z: np.where(z <= 1., z, 2. * np.sqrt(np.maximum(z, 1.)) - 1.)))
    cauchy = staticmethod(_scaled(lambda z: np.log1p(z)))
    arctan = staticmethod(_scaled(lambda z: np.arctan(z)))

    @staticmethod
    def get_ell2_cost_from_residual(f, loss="linear", f_scale=1.):
        """1/2 sum rho(f^2)  (nsol/loss_functions.py:42-46)."""
        f = np.asarray(f, dtype=np.float64)
        return 0.5 * np.sum(LossFunctions.get_loss[loss](f2=f ** 2, f_scale=f_scale))


LossFunctions.get_loss = {
    "linear": LossFunctions.linear,
    "soft_l1": LossFunctions.soft_l1,
    "huber": LossFunctions.huber,
    "cauchy": LossFunctions.cauchy,
    "arctan": LossFunctions.arctan,
}
