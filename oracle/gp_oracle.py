"""CPU oracle for the GP hot path of bbbales2/gp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (gp_b200/) never does: it fails loudly when the CUDA
library is missing.

PARITY STATUS: "parity unpinned" for the Stan-Math rows (a1-a8): the reference stores no golden
output for this path and its arithmetic lives in Stan Math / Eigen / R, none of which is vendored or
installable here (no R, Rcpp, StanHeaders, Eigen in this image).  Those rows restate the published
semantics of Stan Math 2.15-2.17 (the era the reference was written against; no version is pinned
in the reference) at the reference's own call sites.  Rows a9/a10 (derivative kernels, noise-added
solve) ARE pinned: tests/golden/gp_derivs_golden.npz holds the outputs of the reference's own
gp_derivs.py executed in the build container (tests/golden/make_golden.py).

Every function cites the reference file:line (relative to /root/reference) that it restates.
All arithmetic is IEEE float64 through NumPy/SciPy (LAPACK potrf/trtrs via OpenBLAS).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

LOG_TWO_PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------------------
# a4  cov_exp_quad  (models/fit_hyperparameters.stan:19, exact_gp.stan:17, fit_full_gp.stan:18,
#     heteroscedastic.stan:23, westbrook_exact.stan:17, interpolated_gp.stan:10,16)
# Stan Math semantics: K[i,i] = alpha^2 exactly; K[i,j] = alpha^2 * exp(-0.5 * (x_i-x_j)^2 / rho^2)
# for the lower triangle, mirrored into the upper.
# --------------------------------------------------------------------------------------------------
def cov_exp_quad(x, alpha, rho):
    x = np.asarray(x, dtype=np.float64)
    d = x[:, None] - x[None, :]
    K = (alpha * alpha) * np.exp(-0.5 * (d * d) / (rho * rho))
    K = np.tril(K) + np.tril(K, -1).T
    K[np.diag_indices_from(K)] = alpha * alpha
    return K


# a5  diagonal add (fit_hyperparameters.stan:21-24 sigma^2; exact_gp.stan:20-22 1e-10; ...)
def add_diag(K, c):
    K = np.array(K, dtype=np.float64, copy=True)
    K[np.diag_indices_from(K)] += c
    return K


class NotPositiveDefinite(ValueError):
    """Stan Math cholesky_decompose throws std::domain_error when the matrix is not symmetric
    (abs tol 1e-8) or not positive definite; `info` is the LAPACK-style 1-based index of the first
    non-positive pivot (0 when the failure is a symmetry failure)."""

    def __init__(self, msg, info=0):
        super().__init__(msg)
        self.info = info


# a6  cholesky_decompose (fit_hyperparameters.stan:25, exact_gp.stan:23, covariance.cpp:29, ...)
def cholesky_decompose(K):
    K = np.asarray(K, dtype=np.float64)
    if K.ndim != 2 or K.shape[0] != K.shape[1]:
        raise NotPositiveDefinite("cholesky_decompose: matrix is not square")
    if K.size and np.max(np.abs(K - K.T)) > 1e-8:
        raise NotPositiveDefinite("cholesky_decompose: matrix is not symmetric")
    L, info = sla.lapack.dpotrf(K, lower=1, clean=1)
    if info != 0:
        raise NotPositiveDefinite("cholesky_decompose: matrix is not positive definite", int(info))
    return L


def potrf_info(K):
    """LAPACK-convention info of the factorisation (0 ok, k>0 first non-positive pivot)."""
    _, info = sla.lapack.dpotrf(np.asarray(K, dtype=np.float64), lower=1, clean=1)
    return int(info)


# mdivide_left_tri_low (inside multi_normal_cholesky; fit_hyperparameters.stan:31)
def mdivide_left_tri_low(L, b):
    return sla.solve_triangular(L, b, lower=True, check_finite=False)


def mdivide_right_tri_low_T(L, b):
    """solve L^T x = b."""
    return sla.solve_triangular(L, b, lower=True, trans="T", check_finite=False)


# a7  multi_normal_cholesky_lpdf (fit_hyperparameters.stan:31; heteroscedastic_centered.stan:33-34)
# lp = [-0.5 N log(2 pi)] - sum(log L_ii) - 0.5 ||L^-1 (y - mu)||^2 ; the bracket is dropped under `~`.
def multi_normal_cholesky_lpdf(y, mu, L, drop_constants=False):
    y = np.asarray(y, dtype=np.float64)
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64), y.shape)
    z = mdivide_left_tri_low(L, y - mu)
    lp = -np.sum(np.log(np.diag(L))) - 0.5 * float(z @ z)
    if not drop_constants:
        lp -= 0.5 * y.shape[0] * LOG_TWO_PI
    return float(lp)


# a8  non-centred mat-vec f = L z (exact_gp.stan:25, fit_full_gp.stan:26, heteroscedastic.stan:31-32)
def trmv_lower(L, z):
    return np.tril(L) @ np.asarray(z, dtype=np.float64)


# --------------------------------------------------------------------------------------------------
# CS-A: the model block of models/fit_hyperparameters.stan:18-31 as a pure function of
# theta = (alpha, rho, sigma):  K = cov_exp_quad + (sigma^2 + jitter) I ; LML = MVN(y | 0, K).
# --------------------------------------------------------------------------------------------------
def gram_se(x, alpha, rho, diag_add):
    return add_diag(cov_exp_quad(x, alpha, rho), diag_add)


def lml(x, y, alpha, rho, sigma, jitter=0.0, drop_constants=False):
    K = gram_se(x, alpha, rho, sigma * sigma + jitter)
    L = cholesky_decompose(K)
    return multi_normal_cholesky_lpdf(y, 0.0, L, drop_constants)


def lml_grad(x, y, alpha, rho, sigma, jitter=0.0):
    """LML and its gradient w.r.t. (alpha, rho, sigma) -- what Stan's reverse sweep through
    multi_normal_cholesky -> cholesky_decompose -> cov_exp_quad returns:
      dLML/dtheta_p = 0.5 tr((a a^T - K^-1) dK/dtheta_p), a = K^-1 y,
      dK/dalpha = 2 K_se / alpha, dK/drho = K_se * d^2 / rho^3, dK/dsigma = 2 sigma I."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.shape[0]
    Kse = cov_exp_quad(x, alpha, rho)
    K = add_diag(Kse, sigma * sigma + jitter)
    L = cholesky_decompose(K)
    z = mdivide_left_tri_low(L, y)
    a = mdivide_right_tri_low_T(L, z)
    val = -0.5 * n * LOG_TWO_PI - np.sum(np.log(np.diag(L))) - 0.5 * float(z @ z)
    Linv = sla.solve_triangular(L, np.eye(n), lower=True, check_finite=False)
    Kinv = Linv.T @ Linv
    M = np.outer(a, a) - Kinv
    d = x[:, None] - x[None, :]
    g_alpha = 0.5 * np.sum(M * (2.0 * Kse / alpha))
    g_rho = 0.5 * np.sum(M * (Kse * (d * d) / rho ** 3))
    g_sigma = 0.5 * np.trace(M) * 2.0 * sigma
    return float(val), np.array([g_alpha, g_rho, g_sigma])


def lp_fit_hyperparameters(x, y, log_rho, log_alpha, log_sigma):
    """`lp__`-compatible value of models/fit_hyperparameters.stan on the unconstrained scale:
    likelihood with dropped constant (`~`, :31) + gamma(4,4) / normal(0,1) priors with their
    dropped constants (:27-29) + log-Jacobians of the <lower=0> transforms (:13-15)."""
    rho, alpha, sigma = math.exp(log_rho), math.exp(log_alpha), math.exp(log_sigma)
    lp = lml(x, y, alpha, rho, sigma, drop_constants=True)
    lp += 3.0 * math.log(rho) - 4.0 * rho          # gamma(4,4) kernel
    lp += -0.5 * alpha * alpha - 0.5 * sigma * sigma  # normal(0,1) kernels
    lp += log_rho + log_alpha + log_sigma           # Jacobians
    return lp


# --------------------------------------------------------------------------------------------------
# a1-a3  rbf_cov_chol (covariance.cpp:9-47): unit-amplitude SE Gram over the FULL square
# exp(-(xi-xj)^2/(2 l^2)) (:17-21), + 1e-10 on the diagonal (:23-25), Cholesky carried in
# forward-mode duals seeded on l (:13,29); returns L (value) and dLdl (tangent), upper = 0.
# Closed form of the tangent: Ldot = L Phi(L^-1 Kdot L^-T),
# Phi = tril with halved diagonal, Kdot = K_se * d^2 / l^3.
# --------------------------------------------------------------------------------------------------
RBF_JITTER = 1e-10


def rbf_gram_and_tangent(x1, l, jitter=RBF_JITTER):
    x1 = np.asarray(x1, dtype=np.float64)
    d = x1[:, None] - x1[None, :]
    d2 = d * d
    S = np.exp(-d2 / (2.0 * l * l))
    Sdot = S * d2 / (l ** 3)
    S = add_diag(S, jitter)
    return S, Sdot


def phi_lower(A):
    P = np.tril(A)
    P[np.diag_indices_from(P)] *= 0.5
    return P


def rbf_cov_chol(x1, l, jitter=RBF_JITTER):
    S, Sdot = rbf_gram_and_tangent(x1, l, jitter)
    L = cholesky_decompose(S)
    T = sla.solve_triangular(L, Sdot, lower=True, check_finite=False)
    A = sla.solve_triangular(L, T.T, lower=True, check_finite=False).T
    dLdl = L @ phi_lower(A)
    return L, np.tril(dLdl)


def rbf_cov_chol_dual(x1, l, jitter=RBF_JITTER):
    """Literal restatement of covariance.cpp:13-39: an unblocked LLT executed on (value, tangent)
    pairs, i.e. what Eigen's LLT does when instantiated on stan::math::fvar<double>.  O(N^3)
    Python-level work: small N only.  Used to pin the closed form above."""
    S, Sdot = rbf_gram_and_tangent(x1, l, jitter)
    n = S.shape[0]
    Lv = np.zeros((n, n))
    Lt = np.zeros((n, n))
    for j in range(n):
        sv = S[j, j] - Lv[j, :j] @ Lv[j, :j]
        st = Sdot[j, j] - 2.0 * (Lv[j, :j] @ Lt[j, :j])
        if not sv > 0.0:
            raise NotPositiveDefinite("rbf_cov_chol: not positive definite", j + 1)
        dv = math.sqrt(sv)
        dt = 0.5 * st / dv
        Lv[j, j], Lt[j, j] = dv, dt
        if j + 1 < n:
            cv = S[j + 1:, j] - Lv[j + 1:, :j] @ Lv[j, :j]
            ct = Sdot[j + 1:, j] - Lv[j + 1:, :j] @ Lt[j, :j] - Lt[j + 1:, :j] @ Lv[j, :j]
            Lv[j + 1:, j] = cv / dv
            Lt[j + 1:, j] = (ct - Lv[j + 1:, j] * dt) / dv
    return Lv, Lt


# --------------------------------------------------------------------------------------------------
# a9  derivative kernels, API #2 (derivative_kernels.R:39-73): unit amplitude, scalar l,
# element-wise in (tj, tk).  Q = value, R = first derivative, T = second derivative.
# --------------------------------------------------------------------------------------------------
def _e(tj, tk, l):
    tj = np.asarray(tj, dtype=np.float64)
    tk = np.asarray(tk, dtype=np.float64)
    d = tj - tk
    return np.exp(-(d ** 2 / (2.0 * l ** 2))), d


def dk_QQ(tj, tk, l):  # derivative_kernels.R:39-41
    e, _ = _e(tj, tk, l)
    return e


def dk_QR(tj, tk, l):  # :43-45
    e, d = _e(tj, tk, l)
    return (e * d) / l ** 2


def dk_RQ(tj, tk, l):  # :47-49
    return dk_QR(tk, tj, l)


def dk_RR(tj, tk, l):  # :51-53
    e, d = _e(tj, tk, l)
    return e / l ** 2 - (e * d ** 2) / l ** 4


def dk_QT(tj, tk, l):  # :55-57
    e, d = _e(tj, tk, l)
    return -(e / l ** 2) + (e * d ** 2) / l ** 4


def dk_TQ(tj, tk, l):  # :59-61
    return dk_QT(tk, tj, l)


def dk_RT(tj, tk, l):  # :63-65
    e, d = _e(tj, tk, l)
    return (3.0 * e * d) / l ** 4 - (e * d ** 3) / l ** 6


def dk_TR(tj, tk, l):  # :67-69
    return dk_RT(tk, tj, l)


def dk_TT(tj, tk, l):  # :71-73
    e, d = _e(tj, tk, l)
    return (3.0 * e) / l ** 4 - (6.0 * e * d ** 2) / l ** 6 + (e * d ** 4) / l ** 8


DERIV_KERNELS = {"QQ": dk_QQ, "QR": dk_QR, "RQ": dk_RQ, "RR": dk_RR, "QT": dk_QT, "TQ": dk_TQ,
                 "RT": dk_RT, "TR": dk_TR, "TT": dk_TT}


def outer_kernel(name, tj, tk, l):
    """R's outer(tj, tk, FUN = function(a, b) kern(a, b, l)) (pendulum_fit.R:238-240)."""
    tj = np.asarray(tj, dtype=np.float64)
    tk = np.asarray(tk, dtype=np.float64)
    return DERIV_KERNELS[name](tj[:, None], tk[None, :], l)


# --------------------------------------------------------------------------------------------------
# a9  R kernel API #1 (R/kernels.R:19-32), phi = (alpha, rho); outer() semantics.
# RR reproduces the operator-precedence quirk of R/kernels.R:31 (phi1^2 multiplies only the first
# term) unless quirk=False.
# --------------------------------------------------------------------------------------------------
def rk_QQ(x, y, phi):  # R/kernels.R:22-24
    e, _ = _e(np.asarray(x, float)[:, None], np.asarray(y, float)[None, :], phi[1])
    return phi[0] ** 2 * e


def rk_QR(x, y, phi):  # R/kernels.R:26-28
    e, d = _e(np.asarray(x, float)[:, None], np.asarray(y, float)[None, :], phi[1])
    return phi[0] ** 2 * (e * d) / phi[1] ** 2


def rk_RR(x, y, phi, quirk=True):  # R/kernels.R:30-32
    e, d = _e(np.asarray(x, float)[:, None], np.asarray(y, float)[None, :], phi[1])
    second = (e * d ** 2) / phi[1] ** 4
    if not quirk:
        second = phi[0] ** 2 * second
    return phi[0] ** 2 * e / phi[1] ** 2 - second


def rk_QQard(X, Y, phi):  # R/kernels.R:19 ; phi = (alpha, rho[D] or scalar)
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    Y = np.atleast_2d(np.asarray(Y, dtype=np.float64))
    rho = np.broadcast_to(np.asarray(phi[1], dtype=np.float64), (X.shape[1],))
    diff = (X[:, None, :] - Y[None, :, :]) / rho[None, None, :]
    return phi[0] ** 2 * np.exp(-0.5 * np.sum(diff ** 2, axis=2))


# --------------------------------------------------------------------------------------------------
# C2 joint covariance of (y, y', y'') on one grid: blocks {QQ,QR,QT; RQ,RR,RT; TQ,TR,TT} * alpha^2
# + diag(noise_y^2, noise_yp^2, noise_ypp^2) + jitter I (design_notes.Rmd:6-46).
# --------------------------------------------------------------------------------------------------
_JOINT = [["QQ", "QR", "QT"], ["RQ", "RR", "RT"], ["TQ", "TR", "TT"]]


def gram_deriv(t, alpha, rho, noise, jitter, nblocks=3, order0=0):
    """order0 > 0 starts the stack at a higher derivative order (order0 = 1, nblocks = 1 is the
    covdd-only covariance of gpderivs.py:62-81)."""
    t = np.asarray(t, dtype=np.float64)
    n = t.shape[0]
    K = np.empty((nblocks * n, nblocks * n))
    for bi in range(nblocks):
        for bj in range(nblocks):
            K[bi * n:(bi + 1) * n, bj * n:(bj + 1) * n] = alpha ** 2 * outer_kernel(_JOINT[order0 + bi][order0 + bj], t, t, rho)
    for bi in range(nblocks):
        idx = np.arange(bi * n, (bi + 1) * n)
        K[idx, idx] += noise[bi] ** 2 + jitter
    return K


# d/dl of the nine kernels, written out term by term from the closed forms above
# (d e / d l = e d^2 / l^3).  Checked against central differences in tests/test_oracle.py.
def _dl_terms(tj, tk, l):
    e, d = _e(tj, tk, l)
    return e, d


def ddl_QQ(tj, tk, l):
    e, d = _dl_terms(tj, tk, l)
    return e * d ** 2 / l ** 3


def ddl_QR(tj, tk, l):
    e, d = _dl_terms(tj, tk, l)
    return e * d ** 3 / l ** 5 - 2.0 * e * d / l ** 3


def ddl_RR(tj, tk, l):
    e, d = _dl_terms(tj, tk, l)
    return -2.0 * e / l ** 3 + 5.0 * e * d ** 2 / l ** 5 - e * d ** 4 / l ** 7


def ddl_RT(tj, tk, l):
    e, d = _dl_terms(tj, tk, l)
    return -12.0 * e * d / l ** 5 + 9.0 * e * d ** 3 / l ** 7 - e * d ** 5 / l ** 9


def ddl_TT(tj, tk, l):
    e, d = _dl_terms(tj, tk, l)
    return -12.0 * e / l ** 5 + 39.0 * e * d ** 2 / l ** 7 - 14.0 * e * d ** 4 / l ** 9 + e * d ** 6 / l ** 11


DERIV_KERNELS_DL = {"QQ": ddl_QQ, "QR": ddl_QR, "RQ": lambda a, b, l: ddl_QR(b, a, l), "RR": ddl_RR,
                    "QT": lambda a, b, l: -ddl_RR(a, b, l), "TQ": lambda a, b, l: -ddl_RR(b, a, l),
                    "RT": ddl_RT, "TR": lambda a, b, l: ddl_RT(b, a, l), "TT": ddl_TT}


def lml_grad_deriv(t, y, alpha, rho, noise, jitter=0.0, nblocks=None, order0=0):
    """LML of the stacked derivative observations y ~ N(0, gram_deriv(...)) and its gradient with
    respect to (alpha, rho, noise[0..nblocks)): what Stan's reverse sweep returns for the dense
    multi_normal model of gpderivs.py:62-83 (there in the (sf2, l2, s2) parametrisation, see
    gpderivs_log_prob_grad) and for the joint covariance of design_notes.Rmd:25-46."""
    t = np.asarray(t, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    noise = np.atleast_1d(np.asarray(noise, dtype=np.float64))
    if nblocks is None:
        nblocks = noise.shape[0]
    n = t.shape[0]
    N = n * nblocks
    K = gram_deriv(t, alpha, rho, noise, jitter, nblocks, order0)
    L = cholesky_decompose(K)
    z = mdivide_left_tri_low(L, y)
    a = sla.solve_triangular(L, z, lower=True, trans="T", check_finite=False)
    val = -0.5 * N * LOG_TWO_PI - np.sum(np.log(np.diag(L))) - 0.5 * float(z @ z)
    Kinv = sla.cho_solve((L, True), np.eye(N), check_finite=False)
    M = np.outer(a, a) - Kinv
    Kk = np.empty((N, N)); dK = np.empty((N, N))
    for bi in range(nblocks):
        for bj in range(nblocks):
            name = _JOINT[order0 + bi][order0 + bj]
            sl = (slice(bi * n, (bi + 1) * n), slice(bj * n, (bj + 1) * n))
            Kk[sl] = outer_kernel(name, t, t, rho)
            dK[sl] = DERIV_KERNELS_DL[name](t[:, None], t[None, :], rho)
    g = np.empty(2 + nblocks)
    g[0] = 0.5 * float(np.sum(M * (2.0 * alpha * Kk)))
    g[1] = 0.5 * float(np.sum(M * (alpha ** 2 * dK)))
    for b in range(nblocks):
        g[2 + b] = 0.5 * 2.0 * noise[b] * float(np.trace(M[b * n:(b + 1) * n, b * n:(b + 1) * n]))
    return float(val), g


def gpderivs_log_prob_grad(t, dx, sf2, l2, s2):
    """The likelihood term of the reference's Stan model in gpderivs.py:62-83:
    Sigma = sf2 * covdd(t_i, t_j, l2) + s2 I, dx ~ multi_normal(0, Sigma), with
    covdd = 2 exp(-d^2/l2) (l2 - 2 d^2) / l2^2 (gpderivs.py:35-37), i.e. the RR kernel with
    alpha^2 = sf2, l^2 = l2 / 2, sigma^2 = s2.  Returns the log density (with constants) and its
    gradient in (sf2, l2, s2) by the chain rule."""
    alpha, l, sigma = np.sqrt(sf2), np.sqrt(l2 / 2.0), np.sqrt(s2)
    v, g = lml_grad_deriv(t, dx, alpha, l, [sigma], 0.0, 1, 1)
    return v, np.array([g[0] / (2.0 * alpha), g[1] / (4.0 * l), g[2] / (2.0 * sigma)])


# --------------------------------------------------------------------------------------------------
# a10 conditioning
# --------------------------------------------------------------------------------------------------
def condMVN(mean, sigma, dependent_ind, given_ind=None, X_given=None):
    """condMVNorm::condMVN as called at ode_gp_library.R:17,32 (0-based index arrays here):
    condMean = m_d + C D^-1 (X - m_g), condVar = B - C D^-1 C^T with B, C, D the
    dependent/cross/given blocks of sigma."""
    mean = np.asarray(mean, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    dep = np.asarray(dependent_ind)
    if given_ind is None or len(given_ind) == 0:
        return mean[dep].copy(), sigma[np.ix_(dep, dep)].copy()
    giv = np.asarray(given_ind)
    B = sigma[np.ix_(dep, dep)]
    C = sigma[np.ix_(dep, giv)]
    D = sigma[np.ix_(giv, giv)]
    CDinv = np.linalg.solve(D, C.T).T
    cmean = mean[dep] + CDinv @ (np.asarray(X_given, dtype=np.float64) - mean[giv])
    cvar = B - CDinv @ C.T
    return cmean, cvar


def p_dotXn(tn, Xn, phi_n, sigma_n, quirk=True):
    """R/ode_gp_library.R:23-33 with UU/UD/DD = QQ/QR/RR of R/kernels.R."""
    tn = np.asarray(tn, dtype=np.float64)
    n = tn.shape[0]
    UU, UD, DD = rk_QQ(tn, tn, phi_n), rk_QR(tn, tn, phi_n), rk_RR(tn, tn, phi_n, quirk)
    K = np.block([[UU + sigma_n ** 2 * np.eye(n), UD], [UD.T, DD]]) + 1e-6 * np.eye(2 * n)
    return condMVN(np.zeros(2 * n), K, np.arange(n, 2 * n), np.arange(n), Xn)


def p_Xn(tn, Xn, phi_n, sigma_n):
    """R/ode_gp_library.R:3-18."""
    tn = np.asarray(tn, dtype=np.float64)
    n = tn.shape[0]
    UU = rk_QQ(tn, tn, phi_n)
    K = np.block([[UU + sigma_n ** 2 * np.eye(n), UU.T], [UU.T, UU]]) + 1e-6 * np.eye(2 * n)
    return condMVN(np.zeros(2 * n), K, np.arange(n, 2 * n), np.arange(n), Xn)


def p_dotXn_solve(tn, Xn, phi_n, sigma_n, quirk=True):
    """R/ode_gp.R:19-32: mn = RQ (QQ + s^2 I)^-1 Xn ; Kn = RR - RQ (QQ + s^2 I)^-1 QR."""
    tn = np.asarray(tn, dtype=np.float64)
    n = tn.shape[0]
    QQ, QR, RR = rk_QQ(tn, tn, phi_n), rk_QR(tn, tn, phi_n), rk_RR(tn, tn, phi_n, quirk)
    RQ = QR.T
    A = QQ + sigma_n ** 2 * np.eye(n)
    return RQ @ np.linalg.solve(A, np.asarray(Xn, float)), RR - RQ @ np.linalg.solve(A, QR)


def gp_condition(K, KsK, KsKs, y, noise_var, jitter):
    """The shared shape of pendulum_fit.R:242-251 / gp_derivs.py:97-113:
    mu = KsK (K + s^2 I)^-1 y ; cov = KsKs - KsK (K + s^2 I)^-1 KsK^T + jitter I."""
    Kt = add_diag(K, noise_var)
    mu = KsK @ np.linalg.solve(Kt, np.asarray(y, float))
    cov = KsKs - KsK @ np.linalg.solve(Kt, KsK.T) + jitter * np.eye(KsKs.shape[0])
    return mu, cov


def sample_derivs_moments(params, ynoise, ti):
    """pendulum_fit.R:227-251 up to (not including) the mvrnorm draw: params = (l, a, sy)."""
    l, a, sy = params
    K = a ** 2 * outer_kernel("QQ", ti, ti, l)
    KsK = a ** 2 * outer_kernel("RQ", ti, ti, l)
    KsKs = a ** 2 * outer_kernel("RR", ti, ti, l)
    return gp_condition(K, KsK, KsKs, ynoise, sy ** 2, 1e-8)


# --------------------------------------------------------------------------------------------------
# f-1 cubic-Hermite interpolation of tabulated Cholesky factors (covariance.cpp:49-96;
# models/cubic_interpolated_gp.hpp:46-72; formula check cubic_spline_test.R:13-18)
# --------------------------------------------------------------------------------------------------
def _bracket(l, lp):
    lidx = 0
    while lidx < len(lp) - 1:       # covariance.cpp:57-61
        if lp[lidx + 1] >= l:
            break
        lidx += 1
    lidx = min(lidx, len(lp) - 2)
    return lidx


def approx_L(l, lp, Ls, dLdls):
    lidx = _bracket(l, lp)
    x1, x2 = lp[lidx], lp[lidx + 1]
    t = (l - x1) / (x2 - x1)
    y1, y2 = np.tril(Ls[lidx]), np.tril(Ls[lidx + 1])
    k1, k2 = np.tril(dLdls[lidx]), np.tril(dLdls[lidx + 1])
    a = k1 * (x2 - x1) - (y2 - y1)
    b = -k2 * (x2 - x1) + (y2 - y1)
    return (1 - t) * y1 + t * y2 + t * (1 - t) * (a * (1 - t) + b * t)


def approx_Lz(l, lp, Ls, dLdls, z):
    """cubic_interpolated_gp.hpp:46-72: returns (v z, dv/dl z): value and the partial that the
    hand-built precomp_v_vari carries."""
    lidx = _bracket(l, lp)
    x1, x2 = lp[lidx], lp[lidx + 1]
    t = (l - x1) / (x2 - x1)
    dtdl = 1.0 / (x2 - x1)
    y1, y2 = np.asarray(Ls[lidx]), np.asarray(Ls[lidx + 1])
    k1, k2 = np.asarray(dLdls[lidx]), np.asarray(dLdls[lidx + 1])
    a = k1 * (x2 - x1) - (y2 - y1)
    b = -k2 * (x2 - x1) + (y2 - y1)
    v = (1 - t) * y1 + t * y2 + t * (1 - t) * (a * (1 - t) + b * t)
    dvdl = (b * (2 - 3 * t) * t + a * (1 + t * (-4 + 3 * t)) - y1 + y2) * dtdl
    z = np.asarray(z, dtype=np.float64)
    return v @ z, dvdl @ z


# --------------------------------------------------------------------------------------------------
# synthetic workloads of SURVEY 8d (shared by tests and bench so both see the same inputs)
# --------------------------------------------------------------------------------------------------
def synth_xy(n, seed, exact_gp_draw=None):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(0.0, 0.05 * n, size=n))
    f = np.sin(x) + 0.5 * np.sin(3.1 * x)
    y = f + 0.3 * rng.standard_normal(n)
    return x, y


def synth_theta(B, seed):
    rng = np.random.default_rng(seed)
    alpha = np.abs(rng.standard_normal(B)) + 0.1
    rho = rng.gamma(4.0, 1.0 / 4.0, size=B)
    sigma = rng.uniform(0.1, 0.5, size=B)
    return np.stack([alpha, rho, sigma], axis=1)


def lml_grad_lapack(x, y, alpha, rho, sigma, jitter=0.0):
    """Same quantity as lml_grad by the cheapest LAPACK route (dpotrf + dpotri = N^3 flops, all
    BLAS-3, threaded by OpenBLAS) with the trace contractions written as BLAS dot / gemv calls: the
    strongest CPU baseline this oracle can offer; used by bench.py's cpu_baseline and
    --impl reference legs.  Checked against lml_grad in tests/test_oracle.py."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.shape[0]
    d2 = np.subtract.outer(x, x)
    np.square(d2, out=d2)
    E = np.exp(d2 * (-0.5 / (rho * rho)))       # unit-amplitude SE kernel, exact 1 on the diagonal
    K = (alpha * alpha) * E
    K[np.diag_indices_from(K)] = alpha * alpha + sigma * sigma + jitter
    L, info = sla.lapack.dpotrf(K, lower=1, clean=1, overwrite_a=1)
    if info != 0:
        raise NotPositiveDefinite("not positive definite", int(info))
    z = sla.solve_triangular(L, y, lower=True, check_finite=False)
    a = sla.solve_triangular(L, z, lower=True, trans="T", check_finite=False)
    val = -0.5 * n * LOG_TWO_PI - np.sum(np.log(np.diag(L))) - 0.5 * float(z @ z)
    # lower triangle of K^-1; the strict upper triangle stays exactly zero (clean=1 above), so a
    # full-matrix dot with a symmetric matrix S gives sum_{i>=j} Kinv_ij S_ij.
    Kinv, info = sla.lapack.dpotri(L, lower=1, overwrite_c=1)
    kd = np.diag(Kinv).copy()
    ED2 = E * d2
    tr_se = 2.0 * float(np.vdot(Kinv, E)) - float(np.sum(kd))          # tr(Kinv E), E_ii = 1
    tr_d2 = 2.0 * float(np.vdot(Kinv, ED2))                            # diagonal of E*d2 is zero
    s_se = float(a @ (E @ a)) - tr_se
    s_d2 = float(a @ (ED2 @ a)) - tr_d2
    g_alpha = alpha * s_se
    g_rho = 0.5 * alpha * alpha * s_d2 / rho ** 3
    g_sigma = sigma * (float(a @ a) - float(np.sum(kd)))
    return float(val), np.array([g_alpha, g_rho, g_sigma])


# --------------------------------------------------------------------------------------------------
# CS-E  non-centred latent exact GP (models/exact_gp.stan:16-33): lp and gradients, with the gradient
# through the Cholesky taken by the REVERSE-mode adjoint (independent of the forward-mode route the
# CUDA path uses):  Lbar = tril(fbar z^T), Kbar = L^-T Phi(L^T Lbar) L^-1 (symmetrised),
# lbar = sum Kbar * dK/dl.
# --------------------------------------------------------------------------------------------------
def exact_gp_lp_grad(x, y, l, sigma, z, alpha=1.0, jitter=1e-10):
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64); z = np.asarray(z, dtype=np.float64)
    n = x.shape[0]
    Kse = cov_exp_quad(x, alpha, l)
    L = cholesky_decompose(add_diag(Kse, jitter))
    f = L @ z
    r = y - f
    lp = -0.5 * float(z @ z) + 3.0 * math.log(l) - 4.0 * l - n * math.log(sigma) - 0.5 * float(r @ r) / sigma ** 2
    fbar = r / sigma ** 2
    Lbar = np.tril(np.outer(fbar, z))
    P = phi_lower(L.T @ Lbar)
    # Kbar = L^-T P L^-1
    T = sla.solve_triangular(L, P, lower=True, trans="T", check_finite=False)          # L^-T P
    Kbar = sla.solve_triangular(L, T.T, lower=True, trans="T", check_finite=False).T   # (L^-T (L^-T P)^T)^T = L^-T P L^-1
    Kbar = 0.5 * (Kbar + Kbar.T)
    d = x[:, None] - x[None, :]
    g_l = float(np.sum(Kbar * (Kse * d * d / l ** 3))) + 3.0 / l - 4.0
    g_sigma = -n / sigma + float(r @ r) / sigma ** 3
    g_z = -z + L.T @ fbar
    return lp, {"l": g_l, "sigma": g_sigma, "z": g_z, "f": f}


def create_p_dotXnS(Xn_list, mn, Kn, theta, normals, sd_is_variance=True):
    """R/ode_gp_library.R:43-93 with the random normals supplied by the caller (one per call)."""
    X = np.column_stack([np.asarray(c, dtype=np.float64) for c in Xn_list])
    N, D = X.shape
    K_XX = rk_QQard(X, X, theta) + 1e-6 * np.eye(N)
    K_XX_1_mn = np.linalg.solve(K_XX, mn)
    K_XX_1_Kn = np.linalg.solve(K_XX, Kn)
    st = {"i": 1, "A": np.zeros((0, N)), "Kss": np.zeros((0, 0)), "Xs": np.zeros((0, D)), "d": np.zeros(0)}
    it = iter(normals)

    def call(xs_vec):
        xs = np.asarray(xs_vec, dtype=np.float64).reshape(1, D)
        st["A"] = np.vstack([st["A"], rk_QQard(xs, X, theta)])
        kss = rk_QQard(xs, xs, theta)
        if st["Xs"].shape[0]:
            cross = rk_QQard(st["Xs"], xs, theta)
            st["Kss"] = np.block([[st["Kss"], cross], [cross.T, kss]])
        else:
            st["Kss"] = kss
        A = st["A"]
        S = np.linalg.solve(K_XX, A.T)
        m = A @ K_XX_1_mn
        K = st["Kss"] - A @ S + A @ K_XX_1_Kn @ S
        K = (K + K.T) / 2 + 1e-6 * np.eye(K.shape[0])
        i = st["i"]
        if i == 1:
            cmean, cvar = m[0], K[0, 0]
        else:
            cm, cv = condMVN(m, K, [i - 1], list(range(i - 1)), st["d"])
            cmean, cvar = cm[0], cv[0, 0]
        sd = cvar if sd_is_variance else math.sqrt(max(cvar, 0.0))
        dot_xs = cmean + sd * next(it)
        st["i"] = i + 1
        st["Xs"] = np.vstack([st["Xs"], xs]); st["d"] = np.append(st["d"], dot_xs)
        return {"mu": float(cmean), "sigma": float(cvar), "dot_xs": float(dot_xs)}
    return call


# f-3: eigen-basis factor (models/westbrook.stan:2-30; spectral_test.R:6-27)
def approx_L_basis(M, scale, xt, sigma, l):
    x = np.asarray(xt, dtype=np.float64)
    a = 1.0 / (4.0 * scale ** 2)
    b = 1.0 / (2.0 * l ** 2)
    epsilon = math.sqrt(b)
    alpha = math.sqrt(2.0 * a)
    beta = (1.0 + (2.0 * epsilon / alpha) ** 2) ** 0.25
    delta = math.sqrt(alpha ** 2 * (beta ** 2 - 1.0) / 2.0)
    Ht = np.zeros((x.shape[0], M))
    xp = alpha * beta * x
    f = math.sqrt(epsilon ** 2 / (alpha ** 2 + delta ** 2 + epsilon ** 2))
    Ht[:, 0] = math.sqrt(math.sqrt(alpha ** 2 / (alpha ** 2 + delta ** 2 + epsilon ** 2))) * math.sqrt(beta) * np.exp(-delta ** 2 * x * x)
    if M > 1:
        Ht[:, 1] = f * math.sqrt(1.0 / 2) * 2.0 * xp * Ht[:, 0]
    for n in range(3, M + 1):
        Ht[:, n - 1] = (f * math.sqrt(1.0 / (2.0 * (n - 1))) * 2.0 * xp * Ht[:, n - 2]
                        - f ** 2 * math.sqrt(1.0 / (4.0 * (n - 1) * (n - 2))) * 2.0 * (n - 2) * Ht[:, n - 3])
    return sigma * Ht
