"""Generates tests/golden/gp_derivs_golden.npz by EXECUTING THE REFERENCE'S OWN gp_derivs.py
(/root/reference/gp_derivs.py, unmodified, under runpy with a stub matplotlib) in the build
container.  The reference cannot travel to the GPU box, so the vectors are committed; this script
is the provenance.  Run:  python tests/golden/make_golden.py

What is captured: the reference's nine derivative kernels (gp_derivs.py:15-40)
evaluated on its own grid ts = linspace(0, 5, 25) (:58-63), its Gram builder outputs (:76-95),
its pendulum data y (:66) and its noise-added solve mu / cov (:97-113) -- a = l = 1, s = 0.1.
"""
import os
import runpy
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/gp_derivs.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # stub matplotlib (not installed here; the reference only plots with it)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("plot", "show", "legend", "title", "ylim", "imshow", "figure", "xlabel", "ylabel"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    np.random.seed(0)  # the reference draws samples for its plots only; they are not captured
    g = runpy.run_path(REF, run_name="__ref__")
    ts, tts, y = g["ts"], g["tts"], g["y"]
    out = {"ts": ts, "tts": tts, "y": y, "l": g["l"], "a": g["a"], "s": g["s"],
           "K": g["K"], "KsK": g["KsK"], "KsKs": g["KsKs"], "KsKi": g["KsKi"], "KsKsi": g["KsKsi"]}
    # the reference's own kernel closures on the full grid (row = tj, col = tk)
    for name in ("QQ", "QR", "RQ", "RR", "QT", "TQ", "RT", "TR", "TT"):
        f = g[name]
        out["kern_" + name] = np.array([[f(tj, tk) for tk in ts] for tj in tts])
    # the reference's own posterior mean / covariance (its `mu` and second `cov` definitions)
    out["mu_d"] = g["mu"](g["K"], g["KsK"], y)
    out["cov_d"] = g["cov"](g["K"], g["KsK"], g["KsKs"])
    out["mu_i"] = g["mu"](g["K"], g["KsKi"], y)
    out["cov_i"] = g["cov"](g["K"], g["KsKi"], g["KsKsi"])
    np.savez(os.path.join(HERE, "gp_derivs_golden.npz"), **out)
    print("wrote gp_derivs_golden.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
