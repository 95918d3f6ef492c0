"""GPU tests of the block-cyclic Cholesky through the real backend (gpb200_mg_* over the C ABI).
World size = min(2, visible GPUs): with one GPU the whole schedule still runs (every collective
degenerates), which exercises all five CUDA building blocks; with two it runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, pc, q):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gp_b200 import capi
    from gp_b200.block_cyclic import BlockCyclicGP, GpuPanelBackend
    from oracle import gp_oracle as o
    x, y = o.synth_xy(n, 5)
    h = capi.Handle(rank)
    bc = BlockCyclicGP(n, panel_cols=pc, backend=GpuPanelBackend(h, torch.device("cuda", rank)))
    info = bc.factor(x, 1.0, 1.0, 0.3)
    val = bc.lml(y)
    L = bc.gather_factor() if n <= 2048 else None
    q.put((rank, info, val, L))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("n,pc", [(500, 128), (1000, 256), (2000, 512), (4096, 1024)])
def test_block_cyclic_gpu(n, pc):
    from oracle import gp_oracle as o
    world = min(2, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, pc, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x, y = o.synth_xy(n, 5)
    vref = o.lml(x, y, 1.0, 1.0, 0.3)
    Lref = o.cholesky_decompose(o.gram_se(x, 1.0, 1.0, 0.09)) if n <= 2048 else None
    for rank, info, val, L in res:
        assert info == 0
        assert abs(val - vref) <= 1e-9 * abs(vref), (val, vref)
        if L is not None:
            assert np.max(np.abs(L - Lref)) / np.max(np.abs(Lref)) < 1e-9


# ---- round 2: gradient of config 5 and the collectives enqueued from C -------------------------------------------
def _grad_worker(rank, world, port, n, pc, q):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gp_b200 import capi
    from gp_b200.block_cyclic import BlockCyclicGP, GpuPanelBackend
    from oracle import gp_oracle as o
    x, y = o.synth_xy(n, 5)
    h = capi.Handle(rank)
    be = GpuPanelBackend(h, torch.device("cuda", rank)).init_comm()      # NCCL communicator owned by the C handle
    bc = BlockCyclicGP(n, panel_cols=pc, backend=be, keep_all=True)
    info = bc.factor(x, 1.1, 0.9, 0.3)
    val, grad = bc.lml_grad(y)
    q.put((rank, info, val, grad))
    be.close_comm()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("n,pc", [(500, 128), (1000, 256), (2000, 512), (3000, 256)])
def test_block_cyclic_lml_and_gradient_native_collectives(n, pc):
    """models/fit_hyperparameters.stan:19-31 semantics from the distributed factor: LML AND gradient, panel
    broadcasts and all-reduces enqueued from C (gpb200_mg_bcast / _allreduce).  World = min(2, visible GPUs)."""
    from oracle import gp_oracle as o
    world = min(2, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, n, pc, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x, y = o.synth_xy(n, 5)
    rv, rg = o.lml_grad_lapack(x, y, 1.1, 0.9, 0.3)
    for rank, info, val, grad in res:
        assert info == 0
        assert abs(val - rv) <= 1e-9 * abs(rv), (val, rv)
        assert np.max(np.abs(grad - rg)) <= 1e-9 * np.max(np.abs(rg)), (grad, rg)


@pytest.mark.parametrize("n,pc,world", [(1000, 128, 3), (2100, 256, 4), (1500, 512, 8), (900, 256, 2)])
def test_distributed_gradient_partials_add_up_over_emulated_ranks(n, pc, world):
    """The per-rank building blocks of the distributed gradient need no communication: on ONE GPU, run them for every
    rank r of an emulated world in turn on the same (complete) factor and add the partial sums -- the result must be the
    single-GPU gradient.  This exercises every ownership pattern (world 2..8, ragged last panel) without 8 GPUs."""
    from gp_b200 import capi
    from gp_b200.block_cyclic import BlockCyclicGP, GpuPanelBackend
    from oracle import gp_oracle as o
    dev = torch.device("cuda", 0)
    h = capi.Handle(0)
    be = GpuPanelBackend(h, dev).init_comm()
    x, y = o.synth_xy(n, 7)
    alpha, rho, sigma = 0.9, 1.2, 0.25
    bc = BlockCyclicGP(n, panel_cols=pc, backend=be, keep_all=True)
    assert bc.factor(x, alpha, rho, sigma) == 0
    np_, pcc = bc.np_, bc.pc
    Lsq = be.empty(np_, np_)
    for p in range(bc.npanels):
        be.panel_to_square(n, bc.col0(p), bc.ncols(p), bc.panels[p], bc.ld(p), Lsq)
    ypad = be.from_host(np.concatenate([y, np.zeros(np_ - n)]))
    a_tot = torch.zeros(np_, dtype=torch.float64, device=dev)
    qf = 0.0
    Xs = []
    for r in range(world):
        nm = be.my_columns(n, pcc, r, world)
        Xp = be.empty(np_, max(nm, 1)); S = be.empty(pcc, max(nm, 1)); Wd = be.empty(pcc * pcc, 2 * bc.npanels)
        z = be.vector(max(nm, 1)); a = be.vector(np_); s2 = be.vector(2); part = be.vector(8 * np_)
        be.inverse_rows(n, pcc, r, world, Lsq, Xp, S, Wd)
        be.solve_partials(n, pcc, r, world, Lsq, Xp, ypad, z, a, s2, part)
        a_tot += a
        qf += float(s2[0].item())
        logdet = float(s2[1].item())
        Xs.append(Xp)
    sums = np.zeros(3)
    nt = np_ // 128
    theta3 = be.from_host(np.array([alpha, rho, sigma]))
    zero = be.vector(np_)
    for r in range(world):
        partial = be.vector(16 * nt * (nt + 1) // 2)
        s5 = be.vector(5)
        be.trace_partials(n, pcc, r, world, Xs[r], bc.x_dev, zero, theta3, partial, s5[0:3])
        be.quadform_partials(n, r, world, bc.x_dev, a_tot, theta3, partial, s5[3:5])
        v = s5.cpu().numpy()
        sums += np.array([v[0] + v[3], v[1] + v[4], v[2]])
    aa = float((a_tot[:n] ** 2).sum().item())
    lml = -0.5 * n * np.log(2 * np.pi) - logdet - 0.5 * qf
    grad = np.array([alpha * sums[0], 0.5 * alpha ** 2 * sums[1] / rho ** 3, sigma * (aa - sums[2])])
    rv, rg = o.lml_grad_lapack(x, y, alpha, rho, sigma)
    assert abs(lml - rv) <= 1e-9 * abs(rv), (lml, rv)
    assert np.max(np.abs(grad - rg)) <= 1e-9 * np.max(np.abs(rg)), (grad, rg)
    be.close_comm()
    h.close()
