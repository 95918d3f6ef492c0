"""Host-side mirror of R/ode_gp_library.R (and of the older R/ode_gp.R): the conditioning API.

p_Xn / p_dotXn keep the reference signatures (tn, Xn, phi_n, sigma_n) and return what
condMVNorm::condMVN returns, {"condMean", "condVar"} (R/ode_gp_library.R:17,32).  UU/UD/DD, which
the reference never defines, are QQ/QR/RR of R/kernels.R as R/ode_gp.R:5-8,23-26 shows
(SURVEY Appendix A.2).  p_dotXn_mnKn is the R/ode_gp.R:19-32 variant returning {"mn", "Kn"}.
sample_derivs is pendulum_fit.R:227-255.  The joint matrix is assembled by one Gram kernel and
conditioned by the tiled Cholesky; nothing is computed on the host.
"""
from __future__ import annotations

import numpy as np

from . import capi


def condMVN(mean, sigma, dependent_ind, given_ind, X_given, handle=None):
    """condMVNorm::condMVN for the layout the reference uses: given indices 0..ng-1 first, dependent
    indices ng..N-1 after (R/ode_gp_library.R:17: condMVN(m, K, (N+1):(2*N), 1:N, Xn))."""
    ng = len(given_ind)
    N = np.asarray(sigma).shape[0]
    if list(given_ind) != list(range(ng)) or list(dependent_ind) != list(range(ng, N)):
        raise capi.GpB200Error("condMVN: only the reference's block layout (given block first) is supported")
    cm, cv = (handle or capi.default_handle()).cond_mvn(mean, sigma, ng, X_given)
    return {"condMean": cm, "condVar": cv}


def p_dotXn(tn, Xn, phi_n, sigma_n, quirk=True, handle=None):  # R/ode_gp_library.R:23-33
    h = handle or capi.default_handle()
    tn = np.asarray(tn, dtype=np.float64)
    n = tn.shape[0]
    K = h.gram_deriv(tn, float(phi_n[0]), float(phi_n[1]), [float(sigma_n), 0.0], 1e-6, nblocks=2, quirk=quirk)
    cm, cv = h.cond_mvn(np.zeros(2 * n), K, n, Xn)
    return {"condMean": cm, "condVar": cv}


def p_Xn(tn, Xn, phi_n, sigma_n, handle=None):  # R/ode_gp_library.R:3-18
    h = handle or capi.default_handle()
    tn = np.asarray(tn, dtype=np.float64)
    n = tn.shape[0]
    UU = h.gram_outer("QQ", tn, tn, float(phi_n[1]), float(phi_n[0]) ** 2)
    K = np.empty((2 * n, 2 * n), order="F")
    K[:n, :n] = UU + float(sigma_n) ** 2 * np.eye(n)
    K[:n, n:] = UU.T
    K[n:, :n] = UU.T
    K[n:, n:] = UU
    K[np.diag_indices(2 * n)] += 1e-6
    cm, cv = h.cond_mvn(np.zeros(2 * n), K, n, Xn)
    return {"condMean": cm, "condVar": cv}


def p_dotXn_mnKn(tn, Xn, phi_n, sigma_n, quirk=True, handle=None):  # R/ode_gp.R:19-32
    h = handle or capi.default_handle()
    a2, l = float(phi_n[0]) ** 2, float(phi_n[1])
    QQ = h.gram_outer("QQ", tn, tn, l, a2)
    RQ = h.gram_outer("QR", tn, tn, l, a2).T
    RR = h.gram_outer("RR_QUIRK" if quirk else "RR", tn, tn, l, a2)
    mn, Kn = h.gp_condition(QQ, RQ, RR, Xn, float(sigma_n) ** 2, 0.0)
    return {"mn": mn, "Kn": Kn}


def sample_derivs_moments(params, ynoise, ti, handle=None):
    """pendulum_fit.R:227-251: params = (l, a, sy); returns build_mu / build_cov."""
    h = handle or capi.default_handle()
    l, a, sy = (float(v) for v in params)
    K = h.gram_outer("QQ", ti, ti, l, a * a)
    KsK = h.gram_outer("RQ", ti, ti, l, a * a)
    KsKs = h.gram_outer("RR", ti, ti, l, a * a)
    return h.gp_condition(K, KsK, KsKs, ynoise, sy * sy, 1e-8)


def sample_derivs(params, ynoise, ti, rng=None, seed=None, handle=None):
    """pendulum_fit.R:227-255 including the MASS::mvrnorm(1, mu, Sigma) draw (:253): the draw is
    mu + L z with L the GPU Cholesky factor of the posterior covariance.  With `seed` the normals come
    from the device generator (gpb200_mvrnorm: Philox, reproducible per seed); with `rng` (a NumPy
    Generator) they are drawn on the host and only the factorisation and L z run on the GPU."""
    h = handle or capi.default_handle()
    mu, cov = sample_derivs_moments(params, ynoise, ti, handle=h)
    if seed is not None:
        return h.mvrnorm(1, mu, cov, int(seed))
    rng = rng or np.random.default_rng()
    L = h.potrf(cov)
    return mu + h.trmv_lower(L, rng.standard_normal(mu.shape[0]))


def create_p_dotXnS(Xn_list, mn, Kn, theta, rng=None, handle=None):
    """R/ode_gp_library.R:43-93: closure that, given a new state x*, returns the conditional normal
    of the derivative there given the data and every derivative already drawn, then draws from it.

    Differences from the R text, all of them benign: the pre-factorisation of K_XX + 1e-6 I is a GPU
    Cholesky instead of qr() (:55-57); `rnorm(1, mean, condVar)` passes a VARIANCE as sd (:84,
    SURVEY Appendix A.3) -- reproduced by default (sd_is_variance=True on the returned closure).
    Returns the closure; each call returns {"mu", "sigma", "dot_xs"} like the reference (:91).
    """
    h = handle or capi.default_handle()
    rng = rng or np.random.default_rng()
    X = np.column_stack([np.asarray(c, dtype=np.float64) for c in Xn_list])
    N, D = X.shape
    mn = np.asarray(mn, dtype=np.float64)
    Kn = np.asarray(Kn, dtype=np.float64)
    K_XX = h.gram_ard(X, X, float(theta[0]), theta[1])
    L = h.potrf(K_XX + 1e-6 * np.eye(N))
    K_XX_1_mn = h.potrs(L, mn)
    K_XX_1_Kn = h.potrs(L, Kn)
    state = {"i": 1, "K_XsX": np.zeros((0, N)), "K_XsXs": np.zeros((0, 0)), "Xs": np.zeros((0, D)),
             "dot_Xs": np.zeros(0)}

    def p_dotXnS(xs_vec, sd_is_variance=True):
        xs = np.asarray(xs_vec, dtype=np.float64).reshape(1, D)
        st = state
        st["K_XsX"] = np.vstack([st["K_XsX"], h.gram_ard(xs, X, float(theta[0]), theta[1])])
        kss = h.gram_ard(xs, xs, float(theta[0]), theta[1])
        if st["Xs"].shape[0]:
            cross = h.gram_ard(st["Xs"], xs, float(theta[0]), theta[1])
            st["K_XsXs"] = np.block([[st["K_XsXs"], cross], [cross.T, kss]])
        else:
            st["K_XsXs"] = kss
        A = st["K_XsX"]
        S = h.potrs(L, np.asfortranarray(A.T))            # solve(K_XX_qr, t(K_XsX))
        m = A @ K_XX_1_mn
        K = st["K_XsXs"] - A @ S + A @ K_XX_1_Kn @ S
        K = (K + K.T) / 2 + 1e-6 * np.eye(K.shape[0])
        i = st["i"]
        if i == 1:
            cmean, cvar = m[0], K[0, 0]
        else:
            # condMVN(m, K, i, 1:(i-1), c(dot_Xs)): last point given all earlier draws
            cm, cv = h.cond_mvn(m, K, i - 1, st["dot_Xs"])
            cmean, cvar = cm[0], cv[0, 0]
        sd = cvar if sd_is_variance else np.sqrt(max(cvar, 0.0))
        dot_xs = cmean + sd * rng.standard_normal()
        st["i"] = i + 1
        st["Xs"] = np.vstack([st["Xs"], xs])
        st["dot_Xs"] = np.append(st["dot_Xs"], dot_xs)
        return {"mu": float(cmean), "sigma": float(cvar), "dot_xs": float(dot_xs)}

    return p_dotXnS
