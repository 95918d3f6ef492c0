"""The reference's CS-A workflow on the GPU path: fit (rho, alpha, sigma) of models/fit_hyperparameters.stan
to data by maximising lp__ (R/tests.R:13-27 runs NUTS and then keeps the arg-max lp__ draw; here L-BFGS on
the same lp__ and its gradient, both evaluated by libgpb200 through gp_b200.stan_math).

    python examples/fit_hyperparameters.py            # needs a B200
"""
import os
import sys

import numpy as np
from scipy.optimize import minimize

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_b200 import stan_math as sm  # noqa: E402


def fit(x, y, u0=(0.0, 0.0, 0.0), handle=None):
    """Returns (rho, alpha, sigma, lp) at the MAP of the unconstrained posterior."""
    def neg(u):
        lp, g = sm.fit_hyperparameters_lp(x, y, u[0], u[1], u[2], handle=handle)
        return -lp, -g
    r = minimize(neg, np.asarray(u0, dtype=float), jac=True, method="L-BFGS-B", bounds=[(-4, 3)] * 3)
    rho, alpha, sigma = np.exp(r.x)
    return rho, alpha, sigma, -r.fun, r.nit


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    t = np.linspace(0, 10, 400)                      # pendulum-like signal (pendulum_fit.R:18-35 shape)
    y = np.sin(1.3 * t) + 0.15 * rng.standard_normal(t.shape[0])
    rho, alpha, sigma, lp, nit = fit(t, y)
    print("MAP after %d L-BFGS iterations: rho=%.3f alpha=%.3f sigma=%.3f lp__=%.3f" % (nit, rho, alpha, sigma, lp))
