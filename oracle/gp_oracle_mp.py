"""mpmath (50-digit) arbiter for the oracle at N <= 64 -- TEST INFRASTRUCTURE ONLY.
Restates models/fit_hyperparameters.stan:18-31 and the gradient formulas of SURVEY Appendix B in
arbitrary precision so that NumPy, C and CUDA results can be ranked against a value that is not
subject to float64 rounding."""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 50


def _gram(x, alpha, rho, diag_add):
    n = len(x)
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            d = mp.mpf(x[i]) - mp.mpf(x[j])
            K[i, j] = mp.mpf(alpha) ** 2 * mp.exp(-d * d / (2 * mp.mpf(rho) ** 2))
        K[i, i] += diag_add
    return K


def lml_grad(x, y, alpha, rho, sigma, jitter=0.0):
    n = len(x)
    alpha, rho, sigma = mp.mpf(alpha), mp.mpf(rho), mp.mpf(sigma)
    K = _gram(x, alpha, rho, sigma ** 2 + mp.mpf(jitter))
    L = mp.cholesky(K)
    yv = mp.matrix([mp.mpf(v) for v in y])
    z = mp.lu_solve(L, yv)
    Kinv = mp.inverse(K)
    a = Kinv * yv
    lml = -mp.mpf(n) / 2 * mp.log(2 * mp.pi) - sum(mp.log(L[i, i]) for i in range(n)) - (z.T * z)[0] / 2
    g = [mp.mpf(0)] * 3
    for i in range(n):
        for j in range(n):
            d = mp.mpf(x[i]) - mp.mpf(x[j])
            kse = alpha ** 2 * mp.exp(-d * d / (2 * rho ** 2))
            m = a[i] * a[j] - Kinv[i, j]
            g[0] += m * 2 * kse / alpha / 2
            g[1] += m * kse * d * d / rho ** 3 / 2
            if i == j:
                g[2] += m * 2 * sigma / 2
    return float(lml), [float(v) for v in g]
