"""The three CPU baselines of SURVEY 8d / BASELINE.md 4 on the GPU box's host cores (test
infrastructure: uses oracle/):
  (1) single-thread plain-C restatement (Stan-Math loop order): the stand-in for one Stan chain
  (2) NumPy/SciPy + OpenBLAS dpotrf/dpotri on all cores (what bench.py reports as cpu_baseline)
  (3) reference-style process parallelism: P = cores independent single-thread evaluations
      (mclapply / rstan cores, pendulum_fit.R:206,268)
N = 4096 unless given; prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import c_oracle as c  # noqa: E402
from oracle import gp_oracle as o  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    cores = os.cpu_count() or 1
    x, y = o.synth_xy(n, 5)
    th = o.synth_theta(cores, 5)
    out = {"n": n, "cores": cores}
    t0 = time.perf_counter(); r1 = c.lml_grad(x, y, th[0]); out["c_single_thread_s_per_eval"] = round(time.perf_counter() - t0, 3)
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=cores):
        o.lml_grad_lapack(x, y, *th[0])
        t0 = time.perf_counter()
        for b in range(3):
            r2 = o.lml_grad_lapack(x, y, *th[b])
        out["lapack_all_cores_s_per_eval"] = round((time.perf_counter() - t0) / 3, 3)
    t0 = time.perf_counter(); res = c.lml_grad_draws(x, y, th, nthreads=cores); dt = time.perf_counter() - t0
    out["c_process_parallel_evals_per_s"] = round(cores / dt, 4)
    out["c_single_thread_evals_per_s"] = round(1.0 / out["c_single_thread_s_per_eval"], 4)
    out["lapack_all_cores_evals_per_s"] = round(1.0 / out["lapack_all_cores_s_per_eval"], 4)
    ref = o.lml_grad_lapack(x, y, *th[0])
    out["c_vs_lapack_rel_diff"] = float(abs(r1[0] - ref[0]) / abs(ref[0]))
    try:
        out["cpu_model"] = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/cpu_baselines.json", "w"), indent=1)


if __name__ == "__main__":
    main()
