"""``nsol.primal_dual_solver_parameter_study.PrimalDualSolverParameterStudy``
(nsol/primal_dual_solver_parameter_study.py:16-78)."""
import numpy as np

import nsol_b200.primal_dual_solver as pd
from nsol_b200.solver_parameter_study import SolverParameterStudy


class PrimalDualSolverParameterStudy(SolverParameterStudy):

    def __init__(self, solver, observer, dir_output, name="PrimalDual",
                 parameters={"alpha": np.arange(0.01, 0.05, 0.005), "alg_type": ["ALG2", "ALG2_AHMOD", "ALG3"]},
                 reconstruction_info={}, append=False):
        if not isinstance(solver, pd.PrimalDualSolver):
            raise TypeError("solver must be of type 'PrimalDualSolver'")
        SolverParameterStudy.__init__(self, solver=solver, parameters=parameters, observer=observer,
                                      dir_output=dir_output, name=name, reconstruction_info=reconstruction_info,
                                      append=append)

    def _get_fileheader(self):
        return self._header_from_keys(["alpha", "iterations", "x_scale", "L2"])
