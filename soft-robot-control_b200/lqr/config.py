"""iLQRConfig -- the solver options of sofacontrol/lqr/config.py:1-31 (same attribute names and default values),
kept as one table so that the kernel-side struct (srcb200_ilqr_config, include/srcb200.h) can be filled from it."""

# name -> (default, meaning)
_OPTIONS = {
    # outer loop (ilqr.py:54, 109-115)
    "max_iter": (50, "loop runs while nbr_iter <= max_iter"),
    "epsilon": (0.1, "converged when 0 <= J_prev - J < epsilon"),
    "include_input_var_constraint": (True, "penalise u_t - u_{t-1} instead of u_t"),
    "do_linesearch": (True, ""),
    "regularize": (True, ""),
    # line search of the forward pass (ilqr.py:62-87)
    "alpha0": (1., "first step size"),
    "alpha_scaling": (0.5, "step-size reduction"),
    "improv_lb": (1e-4, "accept iff improv_lb < actual/predicted decrease <= improv_ub"),
    "improv_ub": (100, ""),
    "alpha_min": (5e-2, "below this the line search has failed and rho is increased"),
    "counter_limit": (5, "consecutive failures before the search is abandoned"),
    # regularisation schedule of the backward pass (ilqr.py:198-217)
    "rho0": (0., ""),
    "drho0": (0., ""),
    "rho_scaling": (1.5, ""),
    "rho_increase_fp": (10., "added to rho after a failed line search"),
    "rho_max": (1e5, ""),
    "rho_min": (1e-3, ""),
    "state_regularization": (True, "rho enters through B^T (P + rho I) B; False: Q_uu + rho I"),
}


class iLQRConfig:
    def __init__(self):
        for name, (default, _) in _OPTIONS.items():
            setattr(self, name, default)

    def __repr__(self):
        return "iLQRConfig(%s)" % ", ".join("%s=%r" % (k, getattr(self, k)) for k in _OPTIONS)
