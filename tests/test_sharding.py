"""CPU tests of the N>1 path: world_size-2 gloo processes shard independent draws with no data-path
collective and gather 5 doubles per item (gp_b200/sharding.py).  The evaluator is injected (the CPU
oracle) because there is no GPU here; on GPUs the same code runs with the NCCL backend."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from gp_b200.sharding import shard_bounds


def test_shard_bounds_cover_everything_once():
    for total in (0, 1, 5, 256, 257, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans[:-1], spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_eval(x, y, theta, jitter):
    from oracle import gp_oracle as o
    B = theta.shape[0]
    lml = np.empty(B); grad = np.empty((B, 3)); info = np.zeros(B, dtype=np.int32)
    for b in range(B):
        xb = x[b] if x.ndim == 2 else x
        yb = y[b] if y.ndim == 2 else y
        lml[b], grad[b] = o.lml_grad(xb, yb, *theta[b], jitter=jitter)
    return lml, grad, info


def _worker(rank, world, port, per_group, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gp_b200.sharding import lml_grad_draws_sharded
    from oracle import gp_oracle as o
    B, n = 5, 40
    th = o.synth_theta(B, 3)
    if per_group:
        xs, ys = zip(*[o.synth_xy(n, 10 + g) for g in range(B)])
        x, y = np.stack(xs), np.stack(ys)
    else:
        x, y = o.synth_xy(n, 1)
    lml, grad, info = lml_grad_draws_sharded(x, y, th, evaluator=_oracle_eval)
    q.put((rank, lml, grad, info))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("per_group", [False, True])
def test_world_size_2_gloo_matches_single_process(per_group):
    from oracle import gp_oracle as o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, per_group, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B, n = 5, 40
    th = o.synth_theta(B, 3)
    if per_group:
        xs, ys = zip(*[o.synth_xy(n, 10 + g) for g in range(B)])
        x, y = np.stack(xs), np.stack(ys)
    else:
        x, y = o.synth_xy(n, 1)
    ref = _oracle_eval(x, y, th, 0.0)
    for rank, lml, grad, info in res:
        assert np.array_equal(lml, ref[0]) and np.array_equal(grad, ref[1]) and np.array_equal(info, ref[2])
