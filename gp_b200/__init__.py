"""gp_b200: B200-native (sm_100a) drop-in for the GP hot path of bbbales2/gp.

Gram (covariance) build -> Cholesky -> solves -> log marginal likelihood and its hyper-parameter
gradient, as hand-written CUDA behind a C ABI (include/gpb200.h, libgpb200.so).  The Python modules
mirror the reference's R-level interface; they only marshal arguments -- all arithmetic runs on the
GPU and there is no CPU fallback.
"""
from . import capi  # noqa: F401
from .capi import GpB200Error, Handle, NotPositiveDefiniteError, default_handle  # noqa: F401

__version__ = "0.1.0"
