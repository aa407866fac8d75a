"""Factory of the deconvolution solvers / parameter studies of the reference API
(``nsol.deconvolution_solver_parameter_study_interface``), re-pointed at the GPU solvers.

``DeconvolutionSolverStudyInterface`` maps ``reconstruction_type`` x ``tv_solver`` to a configured
solver for  min_x 1/2 ||A x - b||^2 + alpha g(x)  plus the dictionary of measures an Observer evaluates
(nsol/deconvolution_solver_parameter_study_interface.py:46-361):

    TK0L2    TikhonovLinearSolver, B = identity                                   (:217-234)
    TK1L2    TikhonovLinearSolver, B = D                                          (:236-253)
    TVL2     tv_solver "PD":   PrimalDualSolver, prox_f = prox_linear_least_squares,
                               prox_g_conj = prox_tv_conj                         (:257-280)
             tv_solver "ADMM": ADMMLinearSolver                                   (:282-299)
    HuberL2  PrimalDualSolver, prox_g_conj = prox_huber_conj                      (:303-325)

``DeconvolutionParameterStudyInterface`` adds the matching ``*SolverParameterStudy`` (:364-552).
Constructor signatures, method names and error behaviour are the reference's; every solver it
builds runs on the device (nsol_b200.*), there is no CPU path.
"""
import numpy as np

import nsol_b200.admm_linear_solver as admm
import nsol_b200.admm_linear_solver_parameter_study as admmparam
import nsol_b200.observer as Observer
import nsol_b200.primal_dual_solver as pd
import nsol_b200.primal_dual_solver_parameter_study as pdparam
import nsol_b200.tikhonov_linear_solver as tk
import nsol_b200.tikhonov_linear_solver_parameter_study as tkparam
from nsol_b200.loss_functions import LossFunctions as loss_fun
from nsol_b200.prior_measures import PriorMeasures as prior_meas
from nsol_b200.proximal_operators import ProximalOperators as prox
from nsol_b200.similarity_measures import SimilarityMeasures

_TYPES = ("TK0L2", "TK1L2", "TVL2", "HuberL2")


class DeconvolutionSolverStudyInterface(object):

    def __init__(self, A, A_adj, D, D_adj, b, x0, alpha, x_scale, iter_max, iterations, minimizer, measures,
                 reconstruction_type, dimension, L2=8, rho=0.5, x_ref=None, x_ref_mask=None, data_loss="linear",
                 data_loss_scale=1, tv_solver="PD", verbose=0, append=0, dtype=None):
        self._A, self._A_adj, self._D, self._D_adj = A, A_adj, D, D_adj
        self._b, self._x0 = b, x0
        self._alpha = alpha
        self._x_scale = x_scale
        self._iter_max, self._iterations = iter_max, iterations
        self._minimizer = minimizer
        self._measures = measures
        self._reconstruction_type = reconstruction_type
        self._dimension = dimension
        self._L2, self._rho = L2, rho
        self._x_ref, self._x_ref_mask = x_ref, x_ref_mask
        self._data_loss, self._data_loss_scale = data_loss, data_loss_scale
        self._tv_solver = tv_solver
        self._verbose = verbose
        self._append = append            # append to (instead of overwrite) an existing study
        self._dtype = dtype              # additive: "float64" (default) | "float32"
        self._solver = None
        self._measures_dic = None

    # -- solver ------------------------------------------------------------------------------
    def set_up_solver(self):
        if self._reconstruction_type not in _TYPES:
            raise KeyError(self._reconstruction_type)       # the reference's dispatch dict raises the same
        self._solver = getattr(self, "_set_up_solver_" + self._reconstruction_type)()

    def get_solver(self):
        if self._solver is None:
            raise RuntimeError("Run 'set_up_solver' first")
        return self._solver

    def _tikhonov(self, B, B_adj):
        return tk.TikhonovLinearSolver(
            A=self._A, A_adj=self._A_adj, B=B, B_adj=B_adj, b=self._b, alpha=self._alpha, x0=self._x0,
            x_scale=self._x_scale, data_loss=self._data_loss, data_loss_scale=self._data_loss_scale,
            iter_max=self._iter_max, minimizer=self._minimizer, verbose=self._verbose, dtype=self._dtype)

    def _set_up_solver_TK0L2(self):
        ident = lambda x: x.flatten()
        return self._tikhonov(ident, ident)

    def _set_up_solver_TK1L2(self):
        return self._tikhonov(self._D, self._D_adj)

    def _primal_dual(self, prox_g_conj, with_loss):
        kw = dict(data_loss=self._data_loss, data_loss_scale=self._data_loss_scale) if with_loss else {}
        prox_f = lambda x, tau: prox.prox_linear_least_squares(
            x=x, tau=tau, A=self._A, A_adj=self._A_adj, b=self._b, x0=self._x0, iter_max=self._iter_max,
            x_scale=self._x_scale, **kw)
        return pd.PrimalDualSolver(
            prox_f=prox_f, prox_g_conj=prox_g_conj, B=self._D, B_conj=self._D_adj, L2=self._L2, alpha=self._alpha,
            x0=self._x0, iterations=self._iterations, x_scale=self._x_scale, verbose=self._verbose, dtype=self._dtype)

    def _set_up_solver_TVL2(self):
        if self._tv_solver == "PD":
            return self._primal_dual(prox.prox_tv_conj, True)
        if self._tv_solver == "ADMM":
            return admm.ADMMLinearSolver(
                A=self._A, A_adj=self._A_adj, b=self._b, B=self._D, B_adj=self._D_adj, alpha=self._alpha, x0=self._x0,
                x_scale=self._x_scale, data_loss=self._data_loss, data_loss_scale=self._data_loss_scale, rho=self._rho,
                iterations=self._iterations, dimension=self._dimension, iter_max=self._iter_max,
                verbose=self._verbose, dtype=self._dtype)
        raise ValueError("tv_solver must be 'PD' or 'ADMM'")    # (the reference fails with UnboundLocalError here)

    def _set_up_solver_HuberL2(self):
        return self._primal_dual(prox.prox_huber_conj, False)

    # -- measures ----------------------------------------------------------------------------
    def set_up_measures(self):
        if self._x_ref is not None:
            if not isinstance(self._x_ref, np.ndarray):
                raise ValueError("Reference x_ref must be of type 1D np.array")
            if self._x_ref.shape != self._x0.shape:
                raise ValueError("Initial value x0 and reference x_ref arrays must be of same shape")
            if self._x_ref_mask is not None:
                if self._x_ref.shape != self._x_ref_mask.shape:
                    raise ValueError("Reference x_ref and reference mask x_ref_mask arrays must be of same shape")
                indices = np.where(self._x_ref_mask > 0)
            else:
                indices = np.where(self._x_ref != np.inf)
            ref = self._x_ref[indices]
            measures_dic = {m: (lambda x, m=m: SimilarityMeasures.similarity_measures[m](x[indices], ref))
                            for m in self._measures}
        else:
            measures_dic = {}
        reg = {
            "TK0L2": lambda x: prior_meas.zeroth_order_tikhonov(x),
            "TK1L2": lambda x: prior_meas.first_order_tikhonov(x, self._D),
            "TVL2": lambda x: prior_meas.total_variation(x, self._D, self._dimension),
            "HuberL2": lambda x: prior_meas.huber(x, self._D, self._dimension),
        }
        measures_dic["Reg"] = reg[self._reconstruction_type]
        measures_dic["Data"] = lambda x: loss_fun.get_ell2_cost_from_residual(
            self._A(x) - self._b, loss=self._data_loss, f_scale=self._data_loss_scale)
        self._measures_dic = measures_dic

    def get_measures(self):
        if self._measures_dic is None:
            raise RuntimeError("Run 'set_up_measures' first")
        return self._measures_dic


class DeconvolutionParameterStudyInterface(DeconvolutionSolverStudyInterface):

    def __init__(self, A, A_adj, D, D_adj, b, x0, alpha, x_scale, iter_max, iterations, minimizer, measures, dimension,
                 reconstruction_type, dir_output, parameters, name, reconstruction_info, L2=8, rho=0.5, x_ref=None,
                 x_ref_mask=None, data_loss="linear", data_loss_scale=1, tv_solver="PD", verbose=0, append=False,
                 dtype=None):
        DeconvolutionSolverStudyInterface.__init__(
            self, A=A, A_adj=A_adj, D=D, D_adj=D_adj, b=b, x0=x0, alpha=alpha, x_scale=x_scale, iter_max=iter_max,
            iterations=iterations, minimizer=minimizer, measures=measures, reconstruction_type=reconstruction_type,
            dimension=dimension, L2=L2, rho=rho, x_ref=x_ref, x_ref_mask=x_ref_mask, data_loss=data_loss,
            data_loss_scale=data_loss_scale, tv_solver=tv_solver, verbose=verbose, append=append, dtype=dtype)
        self._name = name
        self._parameters = parameters
        self._reconstruction_info = reconstruction_info
        self._dir_output = dir_output
        self._parameter_study = None

    def set_up_parameter_study(self):
        self.set_up_solver()
        self.set_up_measures()
        observer = Observer.Observer()
        observer.set_measures(self._measures_dic)
        if self._reconstruction_type in ("TK0L2", "TK1L2"):
            cls = tkparam.TikhonovLinearSolverParameterStudy
        elif self._reconstruction_type == "TVL2" and self._tv_solver == "ADMM":
            cls = admmparam.ADMMLinearSolverParameterStudy
        else:
            cls = pdparam.PrimalDualSolverParameterStudy
        self._parameter_study = cls(self._solver, observer, dir_output=self._dir_output, parameters=self._parameters,
                                    name=self._name, reconstruction_info=self._reconstruction_info, append=self._append)

    def get_parameter_study(self):
        return self._parameter_study
