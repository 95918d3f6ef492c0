"""Host-side mirror of R/kernels.R (kernel API #1): QQ, QR, RR (x, y, phi) with phi = (alpha, rho)
return the full outer-product matrix, QQard(X, Y, phi) the ARD squared-exponential Gram of two
row-observation matrices.  One GPU kernel launch per matrix instead of R's per-element closures
(R/kernels.R:5,19).

RR reproduces the operator-precedence quirk of R/kernels.R:31 (phi1^2 multiplies only the first
term) by default -- pass quirk=False for the mathematically intended kernel
(derivative_kernels.R:51-53 times alpha^2).
"""
from __future__ import annotations

from . import capi


def QQ(x, y, phi, handle=None):  # R/kernels.R:22-24
    return (handle or capi.default_handle()).gram_outer("QQ", x, y, float(phi[1]), float(phi[0]) ** 2)


def QR(x, y, phi, handle=None):  # R/kernels.R:26-28
    return (handle or capi.default_handle()).gram_outer("QR", x, y, float(phi[1]), float(phi[0]) ** 2)


def RR(x, y, phi, quirk=True, handle=None):  # R/kernels.R:30-32
    return (handle or capi.default_handle()).gram_outer("RR_QUIRK" if quirk else "RR", x, y, float(phi[1]),
                                                         float(phi[0]) ** 2)


def QQard(X, Y, phi, handle=None):  # R/kernels.R:19
    return (handle or capi.default_handle()).gram_ard(X, Y, float(phi[0]), phi[1])
