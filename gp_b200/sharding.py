"""Multi-GPU partitioning of independent evaluations (hyper-parameter draws, per-group GPs).

The reference parallelises exactly this axis with OS processes: mclapply(s_list, ..., mc.cores)
(pendulum_fit.R:268,286) and rstan's chains (pendulum_fit.R:206).  Here: one process per GPU
(torch.distributed), a static contiguous split of the B items, NO collective on the data path --
each rank evaluates its slice with its own handle; only the 5 doubles per item of results are
gathered at the end (all_gather over NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous split of `total` items; the first total % world ranks get one extra."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def lml_grad_draws_sharded(x, y, theta, jitter=0.0, evaluator=None, group=None):
    """Evaluate B = len(theta) independent items across the ranks of the default process group.

    x, y: shared (n,) or per item (B, n).  evaluator(x, y, theta_slice, jitter) -> (lml, grad, info)
    defaults to the GPU handle of this rank's device; tests inject a CPU evaluator to exercise the
    plumbing with the gloo backend.  Returns the full (lml[B], grad[B,3], info[B]) on every rank.
    """
    import torch
    import torch.distributed as dist
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    B = theta.shape[0]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(B, rank, world)
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xs = x[lo:hi] if x.ndim == 2 else x
    ys = y[lo:hi] if y.ndim == 2 else y
    if evaluator is None:
        from . import capi
        dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
        h = capi.default_handle(dev)
        evaluator = lambda a, b, t, j: h.lml_grad_batched(a, b, t, j)  # noqa: E731
    if hi > lo:
        lml, grad, info = evaluator(xs, ys, theta[lo:hi], jitter)
    else:
        lml, grad, info = np.empty(0), np.empty((0, 3)), np.empty(0, dtype=np.int32)
    if world == 1:
        return lml, grad, info
    # fixed-size gather: pad every slice to the largest one
    width = (B + world - 1) // world
    mine = np.zeros((width, 5))
    mine[:hi - lo, 0] = lml
    mine[:hi - lo, 1:4] = grad
    mine[:hi - lo, 4] = info
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(mine).to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    full = np.zeros((B, 5))
    for r in range(world):
        a, b = shard_bounds(B, r, world)
        full[a:b] = out[r].cpu().numpy()[:b - a]
    return full[:, 0].copy(), full[:, 1:4].copy(), full[:, 4].astype(np.int32)
