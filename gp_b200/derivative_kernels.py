"""Host-side mirror of derivative_kernels.R:39-73 (kernel API #2: unit amplitude, scalar l).

Same names, same argument order (tj, tk, l), same element-wise semantics as the R closures -- the
values are computed on the GPU by gpb200_kernel_eval.  `outer(ti, tk, FUN)` (pendulum_fit.R:238-240)
is `outer(name, tj, tk, l)` here and runs the tiled Gram kernel (gpb200_gram_outer) in one call
instead of N*M closure calls.
"""
from __future__ import annotations

from . import capi


def _ev(kind, tj, tk, l, handle=None):
    return (handle or capi.default_handle()).kernel_eval(kind, tj, tk, float(l))


def QQ(tj, tk, l, handle=None):  # derivative_kernels.R:39
    return _ev("QQ", tj, tk, l, handle)


def QR(tj, tk, l, handle=None):  # :43
    return _ev("QR", tj, tk, l, handle)


def RQ(tj, tk, l, handle=None):  # :47
    return _ev("RQ", tj, tk, l, handle)


def RR(tj, tk, l, handle=None):  # :51
    return _ev("RR", tj, tk, l, handle)


def QT(tj, tk, l, handle=None):  # :55
    return _ev("QT", tj, tk, l, handle)


def TQ(tj, tk, l, handle=None):  # :59
    return _ev("TQ", tj, tk, l, handle)


def RT(tj, tk, l, handle=None):  # :63
    return _ev("RT", tj, tk, l, handle)


def TR(tj, tk, l, handle=None):  # :67
    return _ev("TR", tj, tk, l, handle)


def TT(tj, tk, l, handle=None):  # :71
    return _ev("TT", tj, tk, l, handle)


def outer(name, tj, tk, l, amp2=1.0, handle=None):
    """R: amp2 * outer(tj, tk, FUN = function(a, b) <name>(a, b, l))."""
    return (handle or capi.default_handle()).gram_outer(name, tj, tk, float(l), float(amp2))
