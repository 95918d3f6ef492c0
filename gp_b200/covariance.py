"""Host-side mirror of covariance.cpp (the reference's one native plug-in).

rbf_cov_chol(x1, l_) keeps the Rcpp export's signature and return value (covariance.cpp:8-9,41-46):
a dict {"L": N x N, "dLdl": N x N}, column-major, strict upper triangle zero, 1e-10 jitter baked in.
approx_L(l, lp, Ls, dLdls) is covariance.cpp:49-96; approx_Lz is the Stan external function of
models/cubic_interpolated_gp.hpp:38-73 (value and the partial its precomp_v_vari carries).
"""
from __future__ import annotations

from . import capi


def rbf_cov_chol(x1, l_, handle=None):
    L, dLdl = (handle or capi.default_handle()).rbf_cov_chol(x1, float(l_))
    return {"L": L, "dLdl": dLdl}


def approx_L(l, lp, Ls, dLdls, handle=None):
    return (handle or capi.default_handle()).approx_L(float(l), lp, Ls, dLdls)


def approx_Lz(l, lp, Ls, dLdls, z, handle=None):
    return (handle or capi.default_handle()).approx_Lz(float(l), lp, Ls, dLdls, z)


def rbf_cov_chol_grid(x1, lp, handle=None):
    """All P tables of a length-scale grid in one batched GPU call (what models/interpolated_gp.stan:15-21
    and the data block of models/cubic_interpolated_gp.stan:11-12 need): returns (Ls, dLdls)."""
    return (handle or capi.default_handle()).rbf_cov_chol_batched(x1, lp)
