"""ctypes loader for the plain-C oracle (oracle/gp_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgp_oracle.so")
_lib = None
_dp = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(_HERE, "gp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_lml_grad.restype = C.c_int
        _lib.oracle_llt.restype = C.c_int
        _lib.oracle_rbf_cov_chol.restype = C.c_int
        _lib.oracle_deriv_kernel.restype = C.c_double
        _lib.oracle_mvn_chol_lpdf.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def cov_exp_quad(x, alpha, rho):
    x = _f(x); n = x.shape[0]
    K = np.empty((n, n), order="F")
    lib().oracle_cov_exp_quad(C.c_int(n), _p(x), C.c_double(alpha), C.c_double(rho), _p(K))
    return K


def llt(K):
    A = np.array(K, dtype=np.float64, order="F", copy=True)
    info = lib().oracle_llt(C.c_int(A.shape[0]), _p(A))
    return A, info


def lml_grad(x, y, theta, jitter=0.0):
    x = _f(x); y = _f(y); th = _f(theta)
    lml = C.c_double(); g = np.empty(3)
    info = lib().oracle_lml_grad(C.c_int(x.shape[0]), _p(x), _p(y), _p(th), C.c_double(jitter),
                                 C.byref(lml), _p(g), None)
    return lml.value, g, info


def lml_grad_draws(x, y, thetas, jitter=0.0, nthreads=1):
    x = _f(x); y = _f(y); th = _f(thetas); B = th.shape[0]
    out = np.empty((B, 5))
    lib().oracle_lml_grad_draws(C.c_int(x.shape[0]), _p(x), _p(y), C.c_int(B), _p(th), C.c_double(jitter),
                                C.c_int(nthreads), _p(out))
    return out


def rbf_cov_chol(x1, l):
    x1 = _f(x1); n = x1.shape[0]
    L = np.empty((n, n), order="F"); dL = np.empty((n, n), order="F")
    info = lib().oracle_rbf_cov_chol(C.c_int(n), _p(x1), C.c_double(l), _p(L), _p(dL))
    return L, dL, info


KINDS = {"QQ": 0, "QR": 1, "RQ": 2, "RR": 3, "QT": 4, "TQ": 5, "RT": 6, "TR": 7, "TT": 8}


def outer_kernel(name, tj, tk, l, amp2=1.0):
    tj = _f(tj); tk = _f(tk)
    K = np.empty((tj.shape[0], tk.shape[0]), order="F")
    lib().oracle_outer_kernel(C.c_int(KINDS[name]), C.c_int(tj.shape[0]), _p(tj), C.c_int(tk.shape[0]), _p(tk),
                              C.c_double(l), C.c_double(amp2), _p(K))
    return K
