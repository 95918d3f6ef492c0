"""Per-kernel-class event times of one N=4096, B=256 LML+gradient step (no result checks: used to time
experiment builds of the library selected with GPB200_LIB)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gp_b200 import capi  # noqa: E402

n, B = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 256)
dev = torch.device("cuda", 0)
h = capi.Handle(0)
stream = torch.cuda.current_stream(dev)
h.set_stream(stream.cuda_stream); h.set_pointer_mode(True)
rng = np.random.default_rng(5)
x = np.sort(rng.uniform(0, 0.05 * n, n)); y = np.sin(x) + 0.3 * rng.standard_normal(n)
th = np.stack([np.abs(rng.standard_normal(B)) + 0.1, rng.gamma(4.0, 0.25, B), rng.uniform(0.1, 0.5, B)], axis=1)
dx = torch.from_numpy(x).to(dev); dy = torch.from_numpy(y).to(dev); dth = torch.from_numpy(th).to(dev)
lml = torch.empty(B, dtype=torch.float64, device=dev); grad = torch.empty(B, 3, dtype=torch.float64, device=dev)
info = torch.zeros(B, dtype=torch.int32, device=dev)
for _ in range(2):
    h.lml_grad_batched_device(n, B, dx, 0, dy, 0, dth, 0.0, True, lml, grad, info)
torch.cuda.synchronize()
h.set_profiling(True)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(stream)
h.lml_grad_batched_device(n, B, dx, 0, dy, 0, dth, 0.0, True, lml, grad, info)
e1.record(stream); torch.cuda.synchronize()
prof = h.get_profile()
print(json.dumps({"lib": capi.LIB_PATH.split("/")[-1], "ms": round(e0.elapsed_time(e1), 2),
                  "classes_ms": {k: round(v[0], 3) for k, v in prof.items()}}))
