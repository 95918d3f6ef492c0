"""Parity cases the round-1 review found missing:
  * gpb200_lml_grad on JITTER-ONLY covariances (exact_gp.stan:21 1e-10, heteroscedastic.stan:27 1e-9,
    westbrook_exact.stan:21 1e-12; cond(K) 1e8..1e12) against the 50-digit mpmath arbiter, with a bound that is
    stated in terms of the conditioning;
  * BASELINE configs 3 and 4 through the sharded entry point on ALL visible GPUs, >= 8 items per rank checked
    against the oracle on the rank that computed them."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.mark.parametrize("jitter,spacing", [(1e-9, 0.25), (1e-10, 0.2), (1e-12, 0.3)])
def test_lml_grad_jitter_only_against_mpmath(handle, jitter, spacing):
    """sigma = 0: the diagonal carries only the model's jitter constant.  Neither LAPACK, Eigen nor this library can
    return the LML of such a matrix to 1e-9: the attainable accuracy is proportional to cond(K).  Bound asserted, for the
    value and for each gradient component, relative to its own magnitude:  error <= eps * cond(K)
    (the float64 NumPy and LAPACK oracles sit at 2e-3 .. 3e-2 of that bound on these three cases, measured against the
    50-digit arbiter: 4.5e-9 / 1.1e-8 / 2.1e-6 for the value at cond 1e10 / 1.2e11 / 8e12)."""
    from oracle import gp_oracle_mp as m
    n = 48
    x = spacing * np.arange(n) + 0.01 * np.sin(np.arange(n))
    y = np.sin(x) + 0.3 * np.cos(2.3 * x)
    alpha, rho = 1.0, 1.0
    cond = np.linalg.cond(o.gram_se(x, alpha, rho, jitter))
    assert 1e9 < cond < 1e14, cond
    mv, mg = m.lml_grad(x, y, alpha, rho, 0.0, jitter)
    gv, gg = handle.lml_grad(x, y, (alpha, rho, 0.0), jitter=jitter)
    mg = np.asarray(mg)
    bound = EPS * cond
    assert abs(gv - mv) / abs(mv) <= bound, (gv, mv, bound)
    assert np.all(np.abs(gg[:2] - mg[:2]) / np.abs(mg[:2]) <= bound), (gg, mg, bound)
    assert gg[2] == 0.0 and mg[2] == 0.0          # d/d sigma = 2 sigma (...) vanishes at sigma = 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _sharded_worker(rank, world, port, q):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gp_b200.sharding import lml_grad_draws_sharded, shard_bounds
    out = {}
    # C3 shape: one shared (x, y), theta draws sharded; C4 shape: per-group (x_g, y_g, theta_g)
    n3, B3 = 2048, 16 * world
    x, y = o.synth_xy(n3, 3)
    th = o.synth_theta(B3, 3)
    lml, grad, info = lml_grad_draws_sharded(x, y, th)
    lo, hi = shard_bounds(B3, rank, world)
    worst = 0.0
    for b in np.linspace(lo, hi - 1, 8).astype(int):       # 8 of this rank's own items
        rv, rg = o.lml_grad_lapack(x, y, *th[b])
        worst = max(worst, abs(lml[b] - rv) / abs(rv), float(np.max(np.abs(grad[b] - rg)) / np.max(np.abs(rg))))
    out["c3"] = (worst, int(np.abs(info).sum()))
    n4, G = 1024, 16 * world
    xs, ys = zip(*[o.synth_xy(n4, 40 + g) for g in range(G)])
    X = np.stack(xs); Y = np.stack(ys)
    thg = o.synth_theta(G, 4)
    lml, grad, info = lml_grad_draws_sharded(X, Y, thg)
    lo, hi = shard_bounds(G, rank, world)
    worst = 0.0
    for g in np.linspace(lo, hi - 1, 8).astype(int):
        rv, rg = o.lml_grad_lapack(X[g], Y[g], *thg[g])
        worst = max(worst, abs(lml[g] - rv) / abs(rv), float(np.max(np.abs(grad[g] - rg)) / np.max(np.abs(rg))))
    out["c4"] = (worst, int(np.abs(info).sum()))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_configs_every_rank_checks_its_own_items_against_the_oracle():
    world = max(1, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == list(range(world))
    for rank, out in res:
        for cfg in ("c3", "c4"):
            worst, bad = out[cfg]
            assert bad == 0 and worst < 1e-9, (rank, cfg, worst, bad)
