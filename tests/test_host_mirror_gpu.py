"""GPU tests of the host-side mirror of the reference's R-level interface (same names, argument
order and return fields as R/kernels.R, derivative_kernels.R, R/ode_gp_library.R, R/ode_gp.R,
covariance.cpp) -- written the way R/tests.R exercises the reference: deterministic grids
seq(-2, 2, 0.2), exp(t) data, then numeric comparison with the CPU oracle."""
import numpy as np
import pytest

from gp_b200 import capi
from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_derivative_kernels_elementwise_signatures(handle):
    from gp_b200 import derivative_kernels as dk
    tj = np.linspace(-1, 2, 31); tk = np.linspace(0.5, 1.5, 31)
    for name in ("QQ", "QR", "RQ", "RR", "QT", "TQ", "RT", "TR", "TT"):
        got = getattr(dk, name)(tj, tk, 0.8, handle=handle)
        assert relerr(got, o.DERIV_KERNELS[name](tj, tk, 0.8)) < 1e-13
    assert dk.TT(0.0, 0.0, 1.0, handle=handle) == 3.0          # SURVEY 8c spot value
    assert relerr(dk.outer("RQ", tj, tk, 0.8, 2.0, handle=handle), 2.0 * o.outer_kernel("RQ", tj, tk, 0.8)) < 1e-13


def test_r_kernels_api(handle):
    from gp_b200 import kernels as rk
    x = np.arange(-2, 2.0001, 0.2)
    phi = (1.4, 0.7)
    assert relerr(rk.QQ(x, x, phi, handle=handle), o.rk_QQ(x, x, phi)) < 1e-13
    assert relerr(rk.QR(x, x, phi, handle=handle), o.rk_QR(x, x, phi)) < 1e-13
    assert relerr(rk.RR(x, x, phi, handle=handle), o.rk_RR(x, x, phi, True)) < 1e-13
    X = np.stack([x, np.sin(x)], axis=1)
    assert relerr(rk.QQard(X, X, (1.4, [0.7, 1.1]), handle=handle), o.rk_QQard(X, X, (1.4, [0.7, 1.1]))) < 1e-13


def test_p_Xn_p_dotXn_like_R_tests(handle):
    # R/tests.R:5-9,33-64: t = seq(-2, 2, 0.2), data exp(t)
    from gp_b200 import ode_gp_library as lib
    tn = np.arange(-2, 2.0001, 0.2)
    Xn = np.exp(tn)
    phi, sig = (1.2, 1.0), 0.1
    r = lib.p_dotXn(tn, Xn, phi, sig, handle=handle)
    rm, rv = o.p_dotXn(tn, Xn, phi, sig)
    assert relerr(r["condMean"], rm) < 1e-8 and relerr(r["condVar"], rv) < 1e-8
    r = lib.p_Xn(tn, Xn, phi, sig, handle=handle)
    rm, rv = o.p_Xn(tn, Xn, phi, sig)
    assert relerr(r["condMean"], rm) < 1e-8 and relerr(r["condVar"], rv) < 1e-7
    r = lib.p_dotXn_mnKn(tn, Xn, phi, sig, handle=handle)
    mn, Kn = o.p_dotXn_solve(tn, Xn, phi, sig)
    assert relerr(r["mn"], mn) < 1e-8 and relerr(r["Kn"], Kn) < 1e-8
    # derivative of exp(t) is exp(t): the posterior mean tracks it (the eyeball check of R/tests.R:64-76)
    assert np.max(np.abs(r["mn"][3:-3] - Xn[3:-3])) < 0.5


def test_sample_derivs_draw(handle):
    from gp_b200 import ode_gp_library as lib
    ti = np.linspace(0, 10, 120)
    yn = np.sin(ti) + 0.05 * np.random.default_rng(0).standard_normal(120)
    draw = lib.sample_derivs((1.0, 1.2, 0.05), yn, ti, rng=np.random.default_rng(1), handle=handle)
    mu, cov = o.sample_derivs_moments((1.0, 1.2, 0.05), yn, ti)
    L = np.linalg.cholesky(cov)
    ref = mu + L @ np.random.default_rng(1).standard_normal(120)
    assert relerr(draw, ref) < 1e-6
    assert np.max(np.abs(mu[10:-10] - np.cos(ti[10:-10]))) < 0.2


def test_rbf_cov_chol_return_shape(handle):
    from gp_b200 import covariance as cv
    out = cv.rbf_cov_chol(np.arange(20) * 1.0, 0.6, handle=handle)
    assert set(out) == {"L", "dLdl"} and out["L"].shape == (20, 20) and out["L"].flags.f_contiguous
    Lr, dLr = o.rbf_cov_chol(np.arange(20) * 1.0, 0.6)
    assert relerr(out["L"], Lr) < 1e-9 and relerr(out["dLdl"], dLr) < 1e-8


def test_stan_math_mirror(handle):
    from gp_b200 import GpB200Error, NotPositiveDefiniteError, stan_math as sm
    x = np.linspace(0, 10, 100)
    rng = np.random.default_rng(1)
    y = np.sin(x) + 0.2 * rng.standard_normal(100)
    K = sm.cov_exp_quad(x, 1.0, 1.0, 0.04, handle=handle)
    L = sm.cholesky_decompose(K, handle=handle)
    assert relerr(L, o.cholesky_decompose(o.gram_se(x, 1.0, 1.0, 0.04))) < 1e-9
    lp = sm.multi_normal_cholesky_lpdf(y, np.zeros(100), L, handle=handle)
    assert abs(lp - o.lml(x, y, 1.0, 1.0, 0.2)) < 1e-9 * abs(lp)
    z = rng.standard_normal(100)
    assert relerr(sm.multiply_lower_tri(L, z, handle=handle), L @ z) < 1e-12
    Kbad = K.copy(); Kbad[2, 5] += 1e-6
    with pytest.raises(GpB200Error):
        sm.cholesky_decompose(Kbad, handle=handle)
    Kneg = K.copy(); Kneg[50, 50] = -1
    with pytest.raises(NotPositiveDefiniteError):
        sm.cholesky_decompose(Kneg, handle=handle)
    # lp__ on the unconstrained scale and its gradient (what NUTS consumes)
    u = np.log([0.9, 1.2, 0.3])
    lp, g = sm.fit_hyperparameters_lp(x, y, *u, handle=handle)
    assert abs(lp - o.lp_fit_hyperparameters(x, y, *u)) < 1e-9 * abs(lp)
    for k in range(3):
        up, um = u.copy(), u.copy()
        up[k] += 1e-6; um[k] -= 1e-6
        fd = (o.lp_fit_hyperparameters(x, y, *up) - o.lp_fit_hyperparameters(x, y, *um)) / 2e-6
        assert abs(fd - g[k]) < 1e-5 * max(1.0, abs(g[k]))


def test_latent_exact_gp_gradient_through_cholesky(handle):
    # CS-E, models/exact_gp.stan: the CUDA path takes the gradient through the Cholesky with the
    # forward-mode tangent; the oracle with the reverse-mode adjoint -- they must agree
    from gp_b200 import latent_gp
    rng = np.random.default_rng(0)
    n = 200
    x = np.arange(n) * 0.8 + 0.05 * rng.standard_normal(n)
    y = np.sin(x) + 0.1 * rng.standard_normal(n)
    z = rng.standard_normal(n)
    lp, g = latent_gp.exact_gp_log_prob(x, y, 0.9, 0.3, z, handle=handle)
    rlp, rg = o.exact_gp_lp_grad(x, y, 0.9, 0.3, z)
    assert abs(lp - rlp) <= 1e-9 * abs(rlp)
    assert abs(g["l"] - rg["l"]) <= 1e-8 * abs(rg["l"])
    assert abs(g["sigma"] - rg["sigma"]) <= 1e-9 * abs(rg["sigma"])
    assert relerr(g["z"], rg["z"]) < 1e-9 and relerr(g["f"], rg["f"]) < 1e-9
    # amplitude variant (fit_full_gp.stan) and the alpha tangent
    L, dLa = handle.se_chol_tangent(x, 1.3, 0.9, 1e-6, 0)
    hh = 1e-6
    Lp = o.cholesky_decompose(o.gram_se(x, 1.3 + hh, 0.9, 1e-6)); Lm = o.cholesky_decompose(o.gram_se(x, 1.3 - hh, 0.9, 1e-6))
    assert relerr(dLa, (Lp - Lm) / (2 * hh)) < 1e-6
    assert relerr(L, o.cholesky_decompose(o.gram_se(x, 1.3, 0.9, 1e-6))) < 1e-9


@pytest.mark.parametrize("n", [60, 200, 700])
def test_reverse_mode_cholesky_adjoint_matches_oracle_and_forward_mode(handle, n):
    """f-2: gpb200_latent_forward / _backward (ONE pass of the Cholesky adjoint for all parameters) against the oracle's
    NumPy reverse sweep, against the round-1 forward-mode route, and -- for alpha, which the oracle does not differentiate --
    against central differences of the oracle's lp; then heteroscedastic.stan:23-32's two mat-vecs on one L (nvec = 2)."""
    from gp_b200 import latent_gp
    rng = np.random.default_rng(n)
    x = np.arange(n) * 0.8 + 0.05 * rng.standard_normal(n)
    y = np.sin(x) + 0.1 * rng.standard_normal(n)
    z = rng.standard_normal(n)
    lp, g = latent_gp.exact_gp_log_prob(x, y, 0.9, 0.3, z, handle=handle, reverse=True)
    lpf, gf = latent_gp.exact_gp_log_prob(x, y, 0.9, 0.3, z, handle=handle, reverse=False)
    rlp, rg = o.exact_gp_lp_grad(x, y, 0.9, 0.3, z)
    assert abs(lp - rlp) <= 1e-9 * abs(rlp) and abs(lp - lpf) <= 1e-12 * abs(lp)
    assert abs(g["l"] - rg["l"]) <= 1e-8 * abs(rg["l"]) and abs(g["l"] - gf["l"]) <= 1e-8 * abs(gf["l"])
    assert relerr(g["z"], rg["z"]) < 1e-9 and relerr(g["f"], rg["f"]) < 1e-9
    # amplitude (fit_full_gp.stan:18-26, jitter 1e-6 keeps the finite difference meaningful)
    a0, l0, jit = 1.3, 0.9, 1e-6

    def lp_of(alpha):
        L = o.cholesky_decompose(o.gram_se(x, alpha, l0, jit))
        r = y - L @ z
        return -0.5 * float(r @ r) / 0.3 ** 2
    f = handle.latent_forward(x, a0, l0, jit, z)
    fbar = (y - f) / 0.3 ** 2
    (ga, gl), zbar = handle.latent_backward(x, a0, l0, jit, z, fbar)
    hh = 1e-6
    assert abs(ga - (lp_of(a0 + hh) - lp_of(a0 - hh)) / (2 * hh)) <= 2e-6 * max(1.0, abs(ga))
    L = o.cholesky_decompose(o.gram_se(x, a0, l0, jit))
    assert relerr(zbar, L.T @ fbar) < 1e-9 and relerr(f, L @ z) < 1e-9
    # two mat-vecs on one factor: the adjoints add
    z2 = rng.standard_normal(n); fb2 = rng.standard_normal(n)
    (ga1, gl1), _ = handle.latent_backward(x, a0, l0, jit, z2, fb2)
    (gab, glb), zb = handle.latent_backward(x, a0, l0, jit, np.stack([z, z2]), np.stack([fbar, fb2]))
    assert abs(gab - (ga + ga1)) <= 1e-9 * max(abs(ga), abs(ga1)) and abs(glb - (gl + gl1)) <= 1e-9 * max(abs(gl), abs(gl1))
    assert relerr(zb[1], L.T @ fb2) < 1e-9


def test_create_p_dotXnS_sequential_sampler(handle):
    # R/tests.R:78-99 shape: condition on data, then query new states one at a time
    from gp_b200 import ode_gp_library as lib
    tn = np.arange(-2, 2.0001, 0.4)
    Xn = np.exp(tn)
    # alpha = 1: with any other amplitude the R/kernels.R:31 quirk makes Kn indefinite (eig -0.8 here) and
    # the reference's LU-based condMVN silently returns negative variances; the Cholesky-based GPU path
    # raises NotPositiveDefiniteError instead (checked below)
    theta = (1.0, [1.0])
    mn, Kn = o.p_dotXn_solve(tn, Xn, (1.0, 1.0), 0.1)
    normals = np.random.default_rng(3).standard_normal(6)

    class FixedRng:
        def __init__(self, v): self.v = list(v)
        def standard_normal(self): return self.v.pop(0)

    # incremental=True grows the Cholesky factor of the star points' covariance by one row per call (f-4);
    # incremental=False is the literal rebuild-and-re-solve of R/ode_gp_library.R:65-92: same numbers
    f = lib.create_p_dotXnS([Xn], mn, Kn, theta, rng=FixedRng(normals), handle=handle)
    fl = lib.create_p_dotXnS([Xn], mn, Kn, theta, rng=FixedRng(normals), handle=handle, incremental=False)
    fr = o.create_p_dotXnS([Xn], mn, Kn, theta, normals)
    for xs in (0.3, 0.9, 1.7, 2.5, 0.5, 4.0):
        a, c, b = f([xs]), fl([xs]), fr([xs])
        for got in (a, c):
            assert abs(got["mu"] - b["mu"]) <= 1e-6 * max(1.0, abs(b["mu"]))
            assert abs(got["sigma"] - b["sigma"]) <= 1e-6 * max(1e-6, abs(b["sigma"]))
            assert abs(got["dot_xs"] - b["dot_xs"]) <= 1e-6 * max(1.0, abs(b["dot_xs"]))
    from gp_b200 import NotPositiveDefiniteError
    mn_q, Kn_q = o.p_dotXn_solve(tn, Xn, (1.2, 1.0), 0.1)       # quirk-affected, indefinite Kn
    fq = lib.create_p_dotXnS([Xn], mn_q, Kn_q, (1.2, [1.0]), rng=FixedRng(normals), handle=handle)
    fq([0.3])
    with pytest.raises(NotPositiveDefiniteError):
        for xs in (0.9, 1.7, 2.5):
            fq([xs])


def test_rbf_cov_chol_grid_and_interpolation(handle):
    # test_interpolate.R:9 grid: lp = seq(qgamma(.05,4,4), qgamma(.95,4,4), length = 10) ~ [0.34, 1.94]
    from gp_b200 import covariance as cv
    x1 = np.arange(60) * 1.5
    lp = np.linspace(0.3416, 1.938, 10)
    Ls, dLs = cv.rbf_cov_chol_grid(x1, lp, handle=handle)
    for q in (0, 4, 9):
        Lr, dLr = o.rbf_cov_chol(x1, float(lp[q]))
        assert relerr(Ls[q], Lr) < 1e-9 and relerr(dLs[q], dLr) < 1e-8
        single = cv.rbf_cov_chol(x1, float(lp[q]), handle=handle)
        assert np.array_equal(single["L"], Ls[q]) and np.array_equal(single["dLdl"], dLs[q])
    # cubic Hermite interpolation between the tabulated factors approximates the exact factor
    l = 0.9
    approx = cv.approx_L(l, lp, Ls, dLs, handle=handle)
    exact, _ = o.rbf_cov_chol(x1, l)
    assert relerr(approx, exact) < 1e-3


def test_eigen_basis_approx_L(handle):
    # models/westbrook.stan:2-30 + its self-check (:72); westbrook.R:22 uses M = 10, scale ~ 0.25 on x in [-0.5, 0.5]
    from gp_b200 import approx_gp
    x = np.linspace(-0.5, 0.5, 300)
    for M in (1, 2, 10, 20):
        L = approx_gp.approx_L(M, 0.25, x, 1.3, 0.4, handle=handle)
        assert L.shape == (300, M) and relerr(L, o.approx_L_basis(M, 0.25, x, 1.3, 0.4)) < 1e-12
    e10 = approx_gp.approx_error(10, 0.25, x, 1.0, 0.4, handle=handle)
    e20 = approx_gp.approx_error(20, 0.25, x, 1.0, 0.4, handle=handle)
    Lr = o.approx_L_basis(20, 0.25, x, 1.0, 0.4)
    ref20 = np.log10(np.max(np.abs(o.cov_exp_quad(x, 1.0, 0.4) - Lr @ Lr.T)) + 1e-20)
    assert e20 < e10 < 0 and abs(e20 - ref20) < 1e-6


def test_map_fit_recovers_noise_level(handle):
    # the caller of the hot path (CS-A): maximise lp__ of fit_hyperparameters.stan with GPU value+gradient
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("fit_example", os.path.join(os.path.dirname(__file__), "..", "examples",
                                                                              "fit_hyperparameters.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    rng = np.random.default_rng(1)
    t = np.linspace(0, 10, 300)
    y = np.sin(1.3 * t) + 0.15 * rng.standard_normal(300)
    rho, alpha, sigma, lp, nit = mod.fit(t, y, handle=handle)
    assert 0.10 < sigma < 0.20 and 0.5 < rho < 3.0 and 0.3 < alpha < 3.0
    # a stationary point of the oracle's lp__ as well
    u = np.log([rho, alpha, sigma])
    h = 1e-5
    for k in range(3):
        up, um = u.copy(), u.copy(); up[k] += h; um[k] -= h
        fd = (o.lp_fit_hyperparameters(t, y, *up) - o.lp_fit_hyperparameters(t, y, *um)) / (2 * h)
        assert abs(fd) < 5e-2


# ---- f-4: device RNG and mvrnorm (pendulum_fit.R:253, lorenz.Rmd:105, ch2.py:42-45) -----------------
def test_device_normals_match_the_restated_stream(handle):
    from oracle import philox
    z = handle.normal_fill(1234, 100001)
    ref = philox.normals(1234, 100001)
    assert np.max(np.abs(z - ref)) < 1e-13          # same integers; log / sincos differ by ulps at most
    assert np.array_equal(handle.normal_fill(1234, 500, offset=4000), z[4000:4500])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    with pytest.raises(capi.GpB200Error):
        handle.normal_fill(1, 10, offset=3)


@pytest.mark.parametrize("m,ndraws", [(1, 1), (25, 1), (100, 7), (300, 130)])
def test_mvrnorm_is_mean_plus_chol_times_the_stream(handle, m, ndraws):
    from oracle import philox
    rng = np.random.default_rng(m)
    x = np.sort(rng.uniform(0, 5, m))
    S = o.gram_se(x, 1.3, 0.7, 0.05)
    mu = rng.standard_normal(m)
    out = handle.mvrnorm(ndraws, mu, S, seed=99)
    L = o.cholesky_decompose(S)
    Z = philox.normals(99, m * ndraws).reshape(ndraws, m)
    ref = mu[None, :] + Z @ L.T
    got = out[None, :] if ndraws == 1 else out
    assert got.shape == (ndraws, m)
    assert np.max(np.abs(got - ref)) <= 1e-11 * max(1.0, np.max(np.abs(ref)))
    assert not np.array_equal(handle.mvrnorm(ndraws, mu, S, seed=100), out)


def test_mvrnorm_moments_and_errors(handle):
    rng = np.random.default_rng(0)
    m, nd = 48, 20000
    A = rng.standard_normal((m, m))
    S = A @ A.T / m + 0.1 * np.eye(m)
    mu = np.linspace(-1, 1, m)
    X = handle.mvrnorm(nd, mu, S, seed=7)
    assert np.max(np.abs(X.mean(0) - mu)) < 5 * np.sqrt(np.max(np.diag(S)) / nd)
    C = np.cov(X.T)
    assert np.max(np.abs(C - S)) < 0.06 * np.max(np.abs(S))
    with pytest.raises(capi.NotPositiveDefiniteError):
        handle.mvrnorm(2, None, -np.eye(3), seed=1)


def test_sample_derivs_with_device_rng(handle, golden):
    from gp_b200 import ode_gp_library as lib
    t, y = golden["ts"], golden["y"]
    d1 = lib.sample_derivs((1.0, 1.0, 0.1), y, t, seed=5, handle=handle)
    d2 = lib.sample_derivs((1.0, 1.0, 0.1), y, t, seed=5, handle=handle)
    d3 = lib.sample_derivs((1.0, 1.0, 0.1), y, t, seed=6, handle=handle)
    mu, cov = lib.sample_derivs_moments((1.0, 1.0, 0.1), y, t, handle=handle)
    assert np.array_equal(d1, d2) and not np.array_equal(d1, d3)
    # a draw lies within a few posterior sd of the posterior mean
    assert np.max(np.abs(d1 - mu) / np.sqrt(np.diag(cov))) < 6.0
