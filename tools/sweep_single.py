"""SURVEY config 5 on ONE GPU: a single exact GP, N = 1k..32k, theta = (1, 1, 0.3): times Gram +
Cholesky + solve (LML only) and the full LML + gradient, device-resident, CUDA events.  Reports
Cholesky-phase TFLOP/s from the library's per-class event timers."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gp_b200 import capi  # noqa: E402

TILE = 128


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384, 32768]
    dev = torch.device("cuda", 0)
    h = capi.Handle(0)
    stream = torch.cuda.current_stream(dev)
    h.set_stream(stream.cuda_stream)
    h.set_pointer_mode(True)
    out = []
    for n in sizes:
        rng = np.random.default_rng(5)
        x = np.sort(rng.uniform(0, 0.05 * n, n))
        y = np.sin(x) + 0.5 * np.sin(3.1 * x) + 0.3 * rng.standard_normal(n)
        dx = torch.from_numpy(x).to(dev); dy = torch.from_numpy(y).to(dev)
        dth = torch.tensor([[1.0, 1.0, 0.3]], dtype=torch.float64, device=dev)
        lml = torch.empty(1, dtype=torch.float64, device=dev); grad = torch.empty(1, 3, dtype=torch.float64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        rec = {"n": n}
        for want_grad in (0, 1):
            for _ in range(2):
                h.lml_grad_batched_device(n, 1, dx, 0, dy, 0, dth, 0.0, want_grad, lml, grad, info)
            torch.cuda.synchronize()
            reps = 3 if n <= 8192 else 2
            h.set_profiling(True)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                h.lml_grad_batched_device(n, 1, dx, 0, dy, 0, dth, 0.0, want_grad, lml, grad, info)
            e1.record(stream); torch.cuda.synchronize()
            prof = h.get_profile(); h.set_profiling(False)
            ms = e0.elapsed_time(e1) / reps
            flops = n ** 3 / 3.0 if not want_grad else float(n) ** 3
            key = "lml_grad" if want_grad else "lml_only"
            rec[key + "_ms"] = round(ms, 3)
            rec[key + "_tflops"] = round(flops / ms * 1e-9, 2)
            rec[key + "_classes_ms"] = {k: round(v[0] / reps, 3) for k, v in prof.items()}
        rec["lml"] = float(lml.item()); rec["info"] = int(info.item())
        out.append(rec)
        print(json.dumps(rec), flush=True)
    json.dump(out, open("gpurun_out/sweep_single.json", "w"), indent=1)


if __name__ == "__main__":
    main()
