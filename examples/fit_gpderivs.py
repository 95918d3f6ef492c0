"""The reference's gpderivs.py workflow on the GPU path: a GP observed only through noisy first derivatives
(dx = cos(t) + noise, gpderivs.py:8-16), hyper-parameters (sf2, l2, s2) of its embedded Stan program
(gpderivs.py:25-133) fitted by maximising the model block's log density with the half-Cauchy priors of
:62-64 -- L-BFGS on log-parameters, value and gradient from one fused GPU evaluation per step
(gp_b200.gpderivs.log_prob_grad -> gpb200_lml_grad_deriv_batched, order0 = 1).  Then the posterior of the
function itself is drawn with the device RNG (gpb200_mvrnorm), the way the program's generated quantities
draw xh (:85-133).

    python examples/fit_gpderivs.py            # needs a B200
"""
import os
import sys

import numpy as np
from scipy.optimize import minimize

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_b200 import capi, gpderivs  # noqa: E402


def fit(t, dx, u0=(0.0, 0.0, -2.0), handle=None):
    def neg(u):
        p = np.exp(u)
        lp, g = gpderivs.log_prob_grad(t, dx, p[0], p[1], p[2], priors=True, handle=handle)
        return -(lp + u.sum()), -(g * p + 1.0)          # log-Jacobian of the <lower=0> transforms
    r = minimize(neg, np.asarray(u0, dtype=float), jac=True, method="L-BFGS-B", bounds=[(-6, 5)] * 3)
    return np.exp(r.x), -r.fun, r.nit


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    t = np.linspace(0, 10, 200)
    dx = np.cos(t) + 0.15 * rng.standard_normal(t.shape[0])
    h = capi.default_handle()
    (sf2, l2, s2), lp, nit = fit(t, dx, handle=h)
    print("MAP after %d L-BFGS iterations: sf2=%.3f l2=%.3f s2=%.4f lp=%.3f" % (nit, sf2, l2, s2, lp))
    # posterior of x(t) given dx (up to the unidentified constant: condition on x(0) = 0 like gpderivs.py:49-53)
    l = np.sqrt(l2 / 2.0)
    Kdd = h.gram_outer("RR", t, t, l, sf2)
    Kxd = h.gram_outer("QR", t, t, l, sf2)
    Kxx = h.gram_outer("QQ", t, t, l, sf2)
    mu, cov = h.gp_condition(Kdd, Kxd, Kxx, dx, s2, 1e-8)
    draws = h.mvrnorm(5, mu, cov, seed=1, jitter=1e-8)
    err = np.max(np.abs((mu - mu[0]) - np.sin(t)))
    print("posterior mean of x(t) - x(0) vs sin(t): max abs error %.3f; 5 posterior draws, sd of draw spread %.3f" % (
        err, float(np.std(draws - mu[None, :]))))
