"""Host-side mirror of the eigen-basis approximation used by the reference's approximate-GP models:
approx_L(M, scale, xt, sigma, l) of models/westbrook.stan:2-30 (identical copies in
multiple_players.stan, fit_approx_gp.stan, fit_ko_approx.stan, density_gp.stan, diffusion_gp.stan,
diffusion_state_gp.stan) and bH of spectral_test.R:6-27.  approx_error is the self-check of
models/westbrook.stan:72: log10 of the max-abs entry of cov_exp_quad(x, sigma, l) - L L^T."""
from __future__ import annotations

import numpy as np

from . import capi


def approx_L(M, scale, xt, sigma, l, handle=None):
    return (handle or capi.default_handle()).approx_L_basis(int(M), float(scale), xt, float(sigma), float(l))


def approx_error(M, scale, xt, sigma, l, handle=None):
    h = handle or capi.default_handle()
    L = h.approx_L_basis(int(M), float(scale), xt, float(sigma), float(l))
    K = h.gram_se(xt, float(sigma), float(l), 0.0)
    return float(np.log10(np.max(np.abs(K - L @ L.T)) + 1e-20))
