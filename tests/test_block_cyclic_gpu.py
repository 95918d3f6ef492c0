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
