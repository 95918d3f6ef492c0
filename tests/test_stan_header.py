"""stan/gp_lml_stan.hpp EXECUTED: compiled against the functional mock of Stan Math (stan/mock/stan/math.hpp: a
tape-recording `var`, precomputed_gradients, grad), linked against libgpb200.so and run (stan/mock/run.cpp).
Every overload a Stan model instantiates is called -- all-var (models/fit_hyperparameters.stan:12-16 as parameters),
mixed, all-double, the gpderivs.py:62-83 parametrisation and the joint (y, y', y'') model -- the reverse sweep is run
on the mock tape, and values and adjoints are compared with the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import gp_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "stan", "mock", "run_stan_mock")


def build():
    subprocess.check_call(["bash", os.path.join(ROOT, "stan", "mock", "build_mock.sh")], stdout=subprocess.DEVNULL)
    return EXE


def inputs():
    n = 200
    i = np.arange(n)
    x = 0.05 * i + 0.01 * np.sin(1.7 * i)
    y = np.sin(x) + 0.3 * np.cos(5.0 * x)
    t = 10.0 * i / (n - 1)
    dx = np.cos(t) + 0.05 * np.sin(11.0 * t)
    ystack = np.concatenate([np.sin(t) + 0.05 * np.cos(7.0 * t), np.cos(t) + 0.05 * np.sin(9.0 * t),
                             -np.sin(t) + 0.05 * np.cos(13.0 * t)])
    return x, y, t, dx, ystack


def test_stan_header_builds_links_and_fails_loudly_without_a_gpu():
    exe = build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "no usable B200 GPU" in r.stdout


@pytest.mark.gpu
def test_stan_header_values_and_adjoints_match_the_oracle():
    exe = build()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads(r.stdout)
    x, y, t, dx, ystack = inputs()

    def close(a, b, tol=1e-9):
        a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
        return np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-300)

    rv, rg = o.lml_grad(x, y, 1.1, 0.9, 0.3)
    assert close(out["lml_vvv"]["value"], rv) and close(out["lml_vvv"]["adj"], rg)
    assert close(out["lml_dvd"]["value"], rv) and close(out["lml_dvd"]["adj"][1], rg[1])
    assert out["lml_dvd"]["adj"][0] == 0.0 and out["lml_dvd"]["adj"][2] == 0.0      # double arguments get no partial
    assert close(out["lml_ddd"]["value"], rv)
    dv, dg = o.gpderivs_log_prob_grad(t, dx, 1.3, 1.7, 0.04)
    assert close(out["dd_vvv"]["value"], dv) and close(out["dd_vvv"]["adj"], dg)
    assert close(out["dd_ddd"]["value"], dv)
    jv, jg = o.lml_grad_deriv(t, ystack, 1.2, 1.1, [0.2, 0.25, 0.3], 1e-6, 3, 0)
    assert close(out["joint_vvv"]["value"], jv) and close(out["joint_vvv"]["adj"], jg, 1e-8)
    assert close(out["joint_ddd"]["value"], jv)
    assert out["domain_error_on_singular"] is True
