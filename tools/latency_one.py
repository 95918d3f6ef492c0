"""One LML+gradient evaluation of a single matrix, eagerly (no CUDA graph), for an ncu launch list:
python tools/latency_one.py N [B]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from gp_b200 import capi
n = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h = capi.Handle(0)
rng = np.random.default_rng(5)
x = np.sort(rng.uniform(0, 0.05 * n, n)); y = np.sin(x) + 0.3 * rng.standard_normal(n)
th = np.tile(np.array([[1.0, 1.0, 0.3]]), (B, 1))
for _ in range(3):
    lml, grad, info = h.lml_grad_batched(x, y, th)
print(lml[0], grad[0], info[0])
