"""``nsol.admm_linear_solver_parameter_study.ADMMLinearSolverParameterStudy``
(nsol/admm_linear_solver_parameter_study.py:16-85)."""
import numpy as np

import nsol_b200.admm_linear_solver as admm
from nsol_b200.solver_parameter_study import SolverParameterStudy


class ADMMLinearSolverParameterStudy(SolverParameterStudy):

    def __init__(self, solver, observer, dir_output, name="ADMM",
                 parameters={"alpha": np.arange(0.01, 0.05, 0.01), "rho": np.arange(0.1, 1.5, 0.5)},
                 reconstruction_info={}, append=False):
        if not isinstance(solver, admm.ADMMLinearSolver):
            raise TypeError("solver must be of type 'ADMMLinearSolver'")
        SolverParameterStudy.__init__(self, solver=solver, parameters=parameters, observer=observer,
                                      dir_output=dir_output, name=name, reconstruction_info=reconstruction_info,
                                      append=append)

    def _get_fileheader(self):
        return self._header_from_keys(["alpha", "rho", "iterations", "minimizer", "iter_max", "x_scale",
                                       "data_loss", "data_loss_scale", "dimension"])
