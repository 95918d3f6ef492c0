"""BASELINE configs 3 and 4 at their stated multi-GPU scale, through the public sharded entry point
(gp_b200.sharding.lml_grad_draws_sharded: host arrays in, results gathered on every rank, no collective on
the data path):
  C3  N = 2048, 4096 hyper-parameter draws on one shared (x, y), sharded over the GPUs of the box
  C4  256 independent per-group GPs of N = 1024 (own x_g, y_g, theta_g), sharded over the GPUs
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/bench_sharded_configs.py
Wall time around the whole call (barrier on both sides, max over ranks), one JSON line per config on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_b200.sharding import lml_grad_draws_sharded  # noqa: E402


def timed(fn, reps=3):
    fn()  # warm-up: workspace, task lists
    out = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = fn()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(float(t.item()))
    return min(out), res


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    rng = np.random.default_rng(3)
    # C3
    n, B = 2048, 4096
    x = np.sort(rng.uniform(0, 0.05 * n, n)); y = np.sin(x) + 0.3 * rng.standard_normal(n)
    th = np.stack([np.abs(rng.standard_normal(B)) + 0.1, rng.gamma(4.0, 0.25, B), rng.uniform(0.1, 0.5, B)], axis=1)
    s, (lml, grad, info) = timed(lambda: lml_grad_draws_sharded(x, y, th))
    if rank == 0:
        print(json.dumps({"config": "C3 N=2048 x 4096 draws", "gpus": world, "seconds": round(s, 4), "evals_per_s": round(B / s, 1),
                          "tflops": round(B * float(n) ** 3 / s * 1e-12, 1), "all_pd": bool((info == 0).all()),
                          "lml_checksum": float(np.sum(lml))}), flush=True)
    # C4
    n, G = 1024, 256
    X = np.sort(rng.uniform(0, 0.05 * n, (G, n)), axis=1); Y = np.sin(X) + 0.3 * rng.standard_normal((G, n))
    th = np.stack([np.abs(rng.standard_normal(G)) + 0.1, rng.gamma(4.0, 0.25, G), rng.uniform(0.1, 0.5, G)], axis=1)
    s, (lml, grad, info) = timed(lambda: lml_grad_draws_sharded(X, Y, th))
    if rank == 0:
        print(json.dumps({"config": "C4 256 groups of N=1024", "gpus": world, "seconds": round(s, 5), "evals_per_s": round(G / s, 1),
                          "tflops": round(G * float(n) ** 3 / s * 1e-12, 1), "all_pd": bool((info == 0).all()),
                          "lml_checksum": float(np.sum(lml))}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
