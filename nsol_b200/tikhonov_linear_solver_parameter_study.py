"""``nsol.tikhonov_linear_solver_parameter_study.TikhonovLinearSolverParameterStudy``
(nsol/tikhonov_linear_solver_parameter_study.py:16-81)."""
import numpy as np

import nsol_b200.tikhonov_linear_solver as tk
from nsol_b200.solver_parameter_study import SolverParameterStudy


class TikhonovLinearSolverParameterStudy(SolverParameterStudy):

    def __init__(self, solver, observer, dir_output, name="Tikhonov",
                 parameters={"alpha": np.arange(0.01, 0.05, 0.01)}, reconstruction_info={}, append=False):
        if not isinstance(solver, tk.TikhonovLinearSolver):
            raise TypeError("solver must be of type 'TikhonovLinearSolver'")
        SolverParameterStudy.__init__(self, solver=solver, parameters=parameters, observer=observer,
                                      dir_output=dir_output, name=name, reconstruction_info=reconstruction_info,
                                      append=append)

    def _get_fileheader(self):
        return self._header_from_keys(["alpha", "minimizer", "iter_max", "x_scale", "data_loss", "data_loss_scale"])
