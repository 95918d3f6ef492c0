"""The reference's one native entry point, rbf_cov_chol (covariance.cpp:9-47): GPU path through the C ABI
with HOST buffers (what R's .Call sees: inputs in, two N x N matrices out) next to the plain-C restatement of
the reference's dual-number LLT (oracle/gp_oracle.c, single thread -- the reference is single-threaded Eigen
on fvar<double>).  Test infrastructure (uses oracle/).  Prints one JSON line per N.

    python tests/support/bench_rbf_cov_chol.py [N ...]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gp_b200 import capi  # noqa: E402
from oracle import c_oracle as c  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [100, 500, 1000, 2000, 4096]
    h = capi.Handle(0)
    c.build()
    for n in sizes:
        x = np.linspace(0.0, 10.0, n) if n <= 1000 else np.sort(np.random.default_rng(1).uniform(0, 0.05 * n, n))
        l = 0.7 if n <= 1000 else 1.0
        if n > 1000:   # keep the 1e-10-jitter matrix factorable at large N: widen the spacing relative to l
            x = x * 4.0
        h.rbf_cov_chol(x, l)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            L, dL = h.rbf_cov_chol(x, l)
        gpu_ms = (time.perf_counter() - t0) / reps * 1e3
        rec = {"n": n, "gpu_e2e_ms": round(gpu_ms, 3), "flops_5n3_over_3": 5.0 * n ** 3 / 3.0,
               "gpu_tflops_e2e": round(5.0 * n ** 3 / 3.0 / gpu_ms * 1e-9, 3)}
        if n <= 2000:
            t0 = time.perf_counter()
            Lr, dLr, _ = c.rbf_cov_chol(x, l)
            rec["cpu_dual_llt_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
            rec["speedup"] = round(rec["cpu_dual_llt_ms"] / gpu_ms, 1)
            rec["max_abs_diff_L"] = float(np.max(np.abs(L - Lr)))
        print(json.dumps(rec), flush=True)
    h.close()


if __name__ == "__main__":
    main()
