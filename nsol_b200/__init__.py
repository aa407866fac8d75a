"""nsol_b200 -- B200-native (sm_100a) backend for the iterative proximal-solver hot
path of gift-surg/NSoL, behind NSoL's own Python API.

Drop-in use::

    import nsol_b200.linear_operators as LinearOperators
    import nsol_b200.primal_dual_solver as pd
    from nsol_b200.proximal_operators import ProximalOperators as prox

Module, class and method names follow ``nsol`` (primal_dual_solver,
admm_linear_solver, tikhonov_linear_solver, linear_operators, kernels,
proximal_operators, observer, solver_parameter_study, ...).  All arithmetic runs
in hand-written CUDA kernels through the C ABI of ``libnsol_b200.so``
(include/nsol_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
