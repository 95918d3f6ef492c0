"""r/shim.c EXECUTED: compiled against the functional mock of the R C API (r/mock/), linked against
libgpb200.so and driven through the same name lookup .Call does.  The CPU tests cover the build, the
registration table and the error path; the GPU tests run the entry points a reference user calls
(rbf_cov_chol of covariance.cpp:9, condMVN of R/ode_gp_library.R:17,32, the batched LML over draws of
pendulum_fit.R:259-268) and compare them with the oracle."""
import os
import re
import sys

import numpy as np
import pytest

from oracle import gp_oracle as o
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "support"))
import mock_r  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RFILES = ["gpb200.R", "kernels.R", "derivative_kernels.R", "ode_gp_library.R", "ode_gp.R"]


@pytest.fixture(scope="module")
def R():
    return mock_r.MockR()


def test_shim_builds_links_and_registers_every_call_the_r_files_make(R):
    called = {}
    for f in RFILES:
        txt = open(os.path.join(ROOT, "r", "R", f)).read()
        for m in re.finditer(r'\.Call\("(gp_\w+)"', txt):
            called.setdefault(m.group(1), f)
    assert len(called) >= 15
    for name, f in called.items():
        assert R.registered(name) > 0, "%s (used by r/R/%s) is not registered in r/shim.c" % (name, f)


def test_reference_function_names_are_defined_by_the_r_files():
    """derivative_kernels.R:39-73 names QQ..TT(tj, tk, l); R/kernels.R:19-32 QQ/QR/RR(x, y, phi), QQard;
    R/ode_gp_library.R:3-93 p_Xn, p_dotXn, p_dotX, create_p_dotXnS; covariance.cpp rbf_cov_chol, approx_L."""
    def defs(f):
        return set(re.findall(r"^(\w+)\s*(?:<-|=)\s*function", open(os.path.join(ROOT, "r", "R", f)).read(), re.M))
    assert {"QQ", "QR", "RQ", "RR", "QT", "TQ", "RT", "TR", "TT"} <= defs("derivative_kernels.R")
    assert {"QQ", "QR", "RR", "QQard", "mat_to_obs_list", "obs_list_outer", "create_kernel_function"} <= defs("kernels.R")
    assert {"p_Xn", "p_dotXn", "p_dotX", "create_p_dotXnS"} <= defs("ode_gp_library.R")
    assert {"p_Xn", "p_dotXn", "p_dotX", "create_p_dotXnS"} <= defs("ode_gp.R")
    assert {"rbf_cov_chol", "approx_L"} <= defs("gpb200.R")
    for f, sig in (("derivative_kernels.R", r"TT\s*(?:<-|=)\s*function\(tj, tk, l\)"), ("kernels.R", r"RR\s*<-\s*function\(x,\s*y,\s*phi\)"),
                   ("ode_gp_library.R", r"p_dotXn\s*<-\s*function\(tn, Xn, phi_n, sigma_n\)"),
                   ("ode_gp_library.R", r"create_p_dotXnS\s*<-\s*function\(Xn_list, mn, Kn, theta\)"),
                   ("gpb200.R", r"rbf_cov_chol\s*<-\s*function\(x1, l_\)")):
        assert re.search(sig, open(os.path.join(ROOT, "r", "R", f)).read()), (f, sig)


def test_shape_and_type_errors_are_r_errors_before_any_gpu_work(R):
    x = np.linspace(0, 1, 8)
    with pytest.raises(mock_r.RError, match="3 x B matrix"):
        R.call("gp_lml_grad_draws", x, x, np.ones((5, 3)), 0.0)       # draws-by-parameters layout
    with pytest.raises(mock_r.RError, match=r"length\(y\)"):
        R.call("gp_lml_grad_draws", x, x[:5], np.ones((3, 2)), 0.0)   # short y
    with pytest.raises(mock_r.RError, match="must not be NULL"):
        R.call("gp_rbf_cov_chol", None, 1.0)
    with pytest.raises(mock_r.RError, match="Sigma must be a 4 x 4"):
        R.call("gp_mvrnorm", 2, None, np.ones((4, 3)), 1.0)
    with pytest.raises(mock_r.RError, match="X.given"):
        R.call("gp_cond_mvn", None, np.eye(6), 3, np.zeros(2))
    with pytest.raises(mock_r.RError, match="not registered"):
        R.call("gp_no_such_entry", x)


def test_without_a_gpu_the_shim_raises_an_r_error_not_a_fallback(R):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mock_r.RError, match="no usable B200 GPU"):
        R.call("gp_rbf_cov_chol", np.linspace(0, 1, 8), 1.0)


@pytest.mark.gpu
def test_rbf_cov_chol_through_the_shim(R):
    x = np.linspace(0, 10, 100)                       # test_interpolate.R:5-7 grid
    out = R.call("gp_rbf_cov_chol", x, 1.3)
    assert list(out) == ["L", "dLdl"]                 # covariance.cpp:41-46 names
    S, Sdot = o.rbf_gram_and_tangent(x, 1.3)
    EPS = np.finfo(np.float64).eps
    assert np.linalg.norm(out["L"] @ out["L"].T - S) / np.linalg.norm(S) < 50 * 100 * EPS
    assert np.all(np.triu(out["L"], 1) == 0) and np.all(np.triu(out["dLdl"], 1) == 0)
    # jitter-only matrix (cond ~ 1e10): assert the defining identity of the tangent, no worse than 10x the oracle
    Lr, dLr = o.rbf_cov_chol(x, 1.3)
    Res = out["dLdl"] @ out["L"].T + out["L"] @ out["dLdl"].T - Sdot
    Rr = dLr @ Lr.T + Lr @ dLr.T - Sdot
    assert np.linalg.norm(Res) <= 10 * max(np.linalg.norm(Rr), 1e-12 * np.linalg.norm(Sdot))
    # integer storage is coerced like Rcpp's NumericVector does
    xi = np.arange(12, dtype=np.int32)
    out2 = R.call("gp_rbf_cov_chol", xi, 2)
    Lr2, _ = o.rbf_cov_chol(xi.astype(float), 2.0)
    assert np.max(np.abs(out2["L"] - Lr2)) < 1e-9


@pytest.mark.gpu
def test_cond_mvn_and_p_dotxn_matrix_through_the_shim(R):
    tn = np.arange(-2, 2.01, 0.2)                     # R/tests.R:5
    rng = np.random.default_rng(11)
    Xn = np.exp(tn) + 0.05 * rng.standard_normal(tn.shape[0])
    N = tn.shape[0]
    K = R.call("gp_gram_deriv", tn, 1.2, 0.9, 2, np.array([0.1, 0.0]), 1e-6, 1)
    rmean, rvar = o.p_dotXn(tn, Xn, (1.2, 0.9), 0.1, quirk=True)
    got = R.call("gp_cond_mvn", np.zeros(2 * N), K, N, Xn)
    assert list(got) == ["condMean", "condVar"]       # condMVNorm::condMVN names
    assert np.max(np.abs(got["condMean"] - rmean)) < 1e-7 * np.max(np.abs(rmean))
    assert np.max(np.abs(got["condVar"] - rvar)) < 1e-7
    got0 = R.call("gp_cond_mvn", None, K, N, Xn)      # mean = NULL means zeros
    assert np.array_equal(got0["condMean"], got["condMean"])


@pytest.mark.gpu
def test_lml_grad_draws_through_the_shim(R):
    n, B = 300, 5
    x, y = o.synth_xy(n, 3)
    th = o.synth_theta(B, 7)
    out = R.call("gp_lml_grad_draws", x, y, np.asfortranarray(th.T), 0.0)
    assert out["grad"].shape == (3, B) and np.all(out["info"] == 0)
    for b in range(B):
        rv, rg = o.lml_grad(x, y, *th[b])
        assert abs(out["lml"][b] - rv) <= 1e-9 * abs(rv)
        assert np.max(np.abs(out["grad"][:, b] - rg)) <= 1e-9 * np.max(np.abs(rg))


@pytest.mark.gpu
def test_kernels_ard_potrs_mvrnorm_through_the_shim(R):
    rng = np.random.default_rng(5)
    X = rng.integers(0, 5, (9, 2)).astype(np.int32)   # integer matrix from R
    Y = rng.standard_normal((4, 2))
    K = R.call("gp_gram_ard", X, Y, 1.3, np.array([0.7, 1.9]))
    assert np.max(np.abs(K - o.rk_QQard(X.astype(float), Y, (1.3, np.array([0.7, 1.9]))))) < 1e-14
    tj = np.linspace(0, 3, 7)
    v = R.call("gp_kernel_eval", 8, tj, tj[::-1].copy(), 1.0, 0.8)      # TT
    assert np.max(np.abs(v - o.dk_TT(tj, tj[::-1], 0.8))) < 1e-12
    A = rng.standard_normal((20, 20)); S = A @ A.T + 20 * np.eye(20)
    L = R.call("gp_potrf", S)
    Bm = rng.standard_normal((20, 3))
    Xs = R.call("gp_potrs", L, Bm)
    assert np.max(np.abs(S @ Xs - Bm)) < 1e-10
    xv = R.call("gp_potrs", L, Bm[:, 0].copy())
    assert xv.shape == (20,) and np.max(np.abs(S @ xv - Bm[:, 0])) < 1e-10
    draws = R.call("gp_mvrnorm", 4, None, S, 42.0)
    assert draws.shape == (4, 20) and np.all(np.isfinite(draws))
