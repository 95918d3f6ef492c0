"""Generates tests/golden/ch2_golden.npz by executing the reference's own ch2.py
(/root/reference/ch2.py, unmodified, under runpy with a stub matplotlib).  The last block of that
script (ch2.py:55-91) builds SE Gram matrices in the l2 parametrisation eta2 * exp(-d^2 / l2)
(:61-65), conditions a 1000-point GP on 4 observations (:79-83) and Cholesky-factors the posterior
covariance + 1e-10 I (:84).  Captured: inputs, the posterior mean, a 25x25 sub-sample of the
posterior covariance (every 40th point) and the cross-Gram, so that the fixture stays small.
In this repo's parametrisation: alpha^2 = eta2, rho = sqrt(l2 / 2)."""
import os
import runpy
import sys
import types

import numpy as np

REF = "/root/reference/ch2.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot")
    for name in ("plot", "show", "legend", "title", "ylim", "imshow", "figure", "xlabel", "ylabel"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
    np.random.seed(0)
    g = runpy.run_path(REF, run_name="__ref__")
    idx = np.arange(0, 1000, 40)
    out = {"eta2": g["eta2"], "l2": g["l2"], "sigma2": g["sigma2"], "xs": g["xs"], "xd": g["xd"], "f": g["f"],
           "m": g["m"], "idx": idx, "Kt_sub": g["Kt"][np.ix_(idx, idx)], "Ksd": g["Ksd"], "Kdd": g["Kdd"],
           "Kss_sub": g["Kss"][np.ix_(idx, idx)], "L_diag": np.diag(g["L"]), "L_sub": g["L"][np.ix_(idx, idx)]}
    np.savez(os.path.join(HERE, "ch2_golden.npz"), **out)
    print({k: np.shape(v) for k, v in out.items()})
    L = g["L"]; Kt = g["Kt"]
    print("reference backward error of its own Cholesky:", np.linalg.norm(L @ L.T - Kt - 1e-10 * np.eye(1000)) / np.linalg.norm(Kt))


if __name__ == "__main__":
    main()
