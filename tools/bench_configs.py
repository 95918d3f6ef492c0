"""Throughput of the other BASELINE configs on ONE GPU, device-resident, CUDA events:
  C3  N=2048, B=512 theta-draws on a shared (x, y)   (4096 draws over 8 GPUs)
  C4  N=1024, G=32 and G=256 independent groups, each its own (x, y, theta)
  C1  N=100,  B=4096 draws (small-N, launch-latency regime)
Prints evals/s, whole-step TFLOP/s (N^3 per evaluation) and the per-kernel-class time split."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gp_b200 import capi  # noqa: E402


def run(h, stream, dev, n, B, per_item_xy, reps=3):
    rng = np.random.default_rng(3)
    if per_item_xy:
        x = np.sort(rng.uniform(0, 0.05 * n, (B, n)), axis=1)
        y = np.sin(x) + 0.3 * rng.standard_normal((B, n))
    else:
        x = np.sort(rng.uniform(0, 0.05 * n, n)); y = np.sin(x) + 0.3 * rng.standard_normal(n)
    th = np.stack([np.abs(rng.standard_normal(B)) + 0.1, rng.gamma(4.0, 0.25, B), rng.uniform(0.1, 0.5, B)], axis=1)
    dx = torch.from_numpy(x).to(dev); dy = torch.from_numpy(y).to(dev); dth = torch.from_numpy(th).to(dev)
    lml = torch.empty(B, dtype=torch.float64, device=dev); grad = torch.empty(B, 3, dtype=torch.float64, device=dev)
    info = torch.zeros(B, dtype=torch.int32, device=dev)
    st = n if per_item_xy else 0
    for _ in range(2):
        h.lml_grad_batched_device(n, B, dx, st, dy, st, dth, 0.0, True, lml, grad, info)
    torch.cuda.synchronize()
    h.set_profiling(True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        h.lml_grad_batched_device(n, B, dx, st, dy, st, dth, 0.0, True, lml, grad, info)
    e1.record(stream); torch.cuda.synchronize()
    prof = h.get_profile(); h.set_profiling(False)
    ms = e0.elapsed_time(e1) / reps
    assert int(info.abs().sum().item()) == 0
    return {"n": n, "B": B, "per_item_xy": per_item_xy, "ms_per_batch": round(ms, 3), "evals_per_s": round(B / ms * 1e3, 1),
            "tflops": round(B * float(n) ** 3 / ms * 1e-9, 2), "classes_ms": {k: round(v[0] / reps, 3) for k, v in prof.items()}}


def run_deriv(h, stream, dev, n, nblocks, B, reps=3):
    """C2: joint (y, y', y'') covariance, N = nblocks * n, B theta draws on one grid (device-resident)."""
    rng = np.random.default_rng(2)
    t = np.linspace(0, 10, n)
    y = np.concatenate([np.sin(t), np.cos(t), -np.sin(t)][:nblocks]) + 0.2 * rng.standard_normal(n * nblocks)
    th = np.column_stack([rng.uniform(0.7, 1.5, B), rng.uniform(0.8, 1.6, B)] + [rng.uniform(0.1, 0.4, B) for _ in range(nblocks)])
    dt = torch.from_numpy(t).to(dev); dy = torch.from_numpy(y).to(dev); dth = torch.from_numpy(th).to(dev)
    lml = torch.empty(B, dtype=torch.float64, device=dev); grad = torch.empty(B, 2 + nblocks, dtype=torch.float64, device=dev)
    info = torch.zeros(B, dtype=torch.int32, device=dev)

    def call():
        rc = h.lib.gpb200_lml_grad_deriv_batched(h._h, n, 0, nblocks, B, dt.data_ptr(), 0, dy.data_ptr(), 0, dth.data_ptr(), 1e-6, 1,
                                                 lml.data_ptr(), grad.data_ptr(), info.data_ptr())
        assert rc == 0, rc
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    h.set_profiling(True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        call()
    e1.record(stream); torch.cuda.synchronize()
    prof = h.get_profile(); h.set_profiling(False)
    ms = e0.elapsed_time(e1) / reps
    assert int(info.abs().sum().item()) == 0
    N = n * nblocks
    return {"config": "C2 joint derivative covariance", "n_grid": n, "nblocks": nblocks, "N": N, "B": B, "ms_per_batch": round(ms, 3),
            "evals_per_s": round(B / ms * 1e3, 1), "tflops": round(B * float(N) ** 3 / ms * 1e-9, 2),
            "classes_ms": {k: round(v[0] / reps, 3) for k, v in prof.items()}}


def main():
    dev = torch.device("cuda", 0)
    h = capi.Handle(0)
    stream = torch.cuda.current_stream(dev)
    h.set_stream(stream.cuda_stream); h.set_pointer_mode(True)
    out = []
    for n, B, per in [(2048, 512, False), (1024, 256, True), (1024, 32, True), (512, 1024, False), (100, 4096, False), (100, 1, False)]:
        r = run(h, stream, dev, n, B, per)
        out.append(r); print(json.dumps(r), flush=True)
    r = run_deriv(h, stream, dev, 512, 3, 256)
    out.append(r); print(json.dumps(r), flush=True)
    json.dump(out, open("gpurun_out/bench_configs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
