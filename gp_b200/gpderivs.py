"""Host-side mirror of the Stan program embedded in the reference's gpderivs.py:25-133: a GP observed
only through noisy first derivatives.

Same function names and the same (sf2, l2, s2) parametrisation as the reference's `functions` block
(gpderivs.py:26-38; l2 = 2 l^2 of derivative_kernels.R, sf2 = alpha^2, s2 = sigma^2).  `log_prob_grad`
is the likelihood statement of its model block (gpderivs.py:66-83: Sigma = sf2 * covdd + s2 I,
dx ~ multi_normal(0, Sigma)) together with the gradient Stan's reverse sweep would deliver -- one
fused Gram -> Cholesky -> inverse -> trace evaluation on the GPU (gpb200_lml_grad_deriv_batched with
order0 = 1, nblocks = 1).  The Cauchy priors of :62-64 are added by `log_prob_grad(..., priors=True)`.
"""
from __future__ import annotations

import math

import numpy as np

from . import capi


def _l(l2):
    return math.sqrt(float(l2) / 2.0)


def cov(ti, tj, l2, handle=None):  # gpderivs.py:27-29
    return (handle or capi.default_handle()).kernel_eval("QQ", ti, tj, _l(l2))


def covd(ti, tj, l2, handle=None):  # gpderivs.py:31-33  (= QR of derivative_kernels.R:43)
    return (handle or capi.default_handle()).kernel_eval("QR", ti, tj, _l(l2))


def covdd(ti, tj, l2, handle=None):  # gpderivs.py:35-37  (= RR of derivative_kernels.R:51)
    return (handle or capi.default_handle()).kernel_eval("RR", ti, tj, _l(l2))


def log_prob_grad(t, dx, sf2, l2, s2, priors=False, handle=None):
    """Log density of the model block and its gradient with respect to (sf2, l2, s2)."""
    h = handle or capi.default_handle()
    alpha, l, sigma = math.sqrt(sf2), _l(l2), math.sqrt(s2)
    lml, g, info = h.lml_grad_deriv_batched(t, dx, [[alpha, l, sigma]], 0.0, order0=1, nblocks=1)
    if info[0]:
        raise capi.NotPositiveDefiniteError("gpderivs.log_prob_grad", int(info[0]))
    lp = float(lml[0])
    grad = np.array([g[0, 0] / (2.0 * alpha), g[0, 1] / (4.0 * l), g[0, 2] / (2.0 * sigma)])
    if priors:  # half-Cauchy(0, 5), (0, 40), (0, 5) on the positive parameters, up to constants
        for i, (v, s) in enumerate(((sf2, 5.0), (l2, 40.0), (s2, 5.0))):
            lp += -math.log1p((v / s) ** 2)
            grad[i] += -2.0 * v / (s * s + v * v)
    return lp, grad


def joint_lml_grad(t, y_stack, alpha, rho, noise, jitter=1e-6, handle=None):
    """LML + gradient for the joint (y, y', y'') covariance of design_notes.Rmd:25-46 (config C2):
    y_stack = c(y, yp, ypp) on the grid t, noise = per-block sd.  Returns (lml, grad[2 + nblocks])."""
    h = handle or capi.default_handle()
    noise = np.atleast_1d(np.asarray(noise, dtype=np.float64))
    th = np.concatenate([[alpha, rho], noise])[None, :]
    lml, g, info = h.lml_grad_deriv_batched(t, y_stack, th, jitter, order0=0, nblocks=noise.shape[0])
    if info[0]:
        raise capi.NotPositiveDefiniteError("gpderivs.joint_lml_grad", int(info[0]))
    return float(lml[0]), g[0]
