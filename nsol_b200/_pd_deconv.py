"""Primal-dual deconvolution (prox_f = prox_linear_least_squares): SURVEY.md 8(f) row 1.

Not built yet: the denoising prox maps (prox_ell1/ell2_denoising) are the BASELINE
configurations; the deconvolution configurations run through ADMMLinearSolver /
TikhonovLinearSolver.  The call fails loudly instead of falling back to the CPU."""


def run_pd_deconvolution(solver, cfg):
    raise TypeError("PrimalDualSolver with prox_f = prox_linear_least_squares (primal-dual deconvolution) is not "
                    "implemented in the CUDA backend yet; use ADMMLinearSolver (tv_solver='ADMM') for TV-L2 "
                    "deconvolution. There is no CPU fallback.")
