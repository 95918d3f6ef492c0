"""CPU tests of the block-cyclic multi-GPU Cholesky SCHEDULE (gp_b200/block_cyclic.py): ownership,
look-ahead ordering, receive buffers and collectives, run as world_size-2 (and 3) gloo processes with
a NumPy stand-in for the five per-rank GPU calls.  The same schedule drives the CUDA backend on GPUs
(tests/test_block_cyclic_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

TILE = 128


class NumpyPanelBackend:
    """NumPy restatement of the gpb200_mg_* calls on torch CPU tensors (test infrastructure)."""

    def empty(self, rows, cols):
        return torch.zeros((cols, rows), dtype=torch.float64)

    def vector(self, n, zero=True):
        return torch.zeros(n, dtype=torch.float64)

    def from_host(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))

    def info_scalar(self):
        return torch.zeros(1, dtype=torch.int32)

    def to_host(self, t):
        return t.numpy()

    @staticmethod
    def _v(P, ldp, ncols):
        return P.numpy().reshape(-1)[:ldp * ncols].reshape(ncols, ldp).T   # rows x ncols view

    def gram_panel(self, n, x, alpha, rho, diag_add, col0, ncols, P, ldp):
        x = x.numpy()
        npad = col0 + ldp
        V = self._v(P, ldp, ncols)
        r = np.arange(col0, npad); c = np.arange(col0, col0 + ncols)
        xr = np.where(r < n, x[np.minimum(r, n - 1)], 0.0); xc = np.where(c < n, x[np.minimum(c, n - 1)], 0.0)
        K = alpha ** 2 * np.exp(-0.5 * (xr[:, None] - xc[None, :]) ** 2 / rho ** 2)
        valid = (r[:, None] < n) & (c[None, :] < n)
        K = np.where(valid, K, 0.0)
        eq = r[:, None] == c[None, :]
        K[eq & valid] = alpha ** 2 + diag_add
        K[eq & ~valid] = 1.0
        V[:, :] = K

    def panel_factor(self, n, col0, ncols, P, ldp, info):
        V = self._v(P, ldp, ncols)
        A11 = np.tril(V[:ncols, :]) + np.tril(V[:ncols, :], -1).T
        L11 = np.linalg.cholesky(A11)
        V[:ncols, :] = L11
        if ldp > ncols:
            V[ncols:, :] = np.linalg.solve(L11, V[ncols:, :].T).T

    def panel_update(self, n, pcol0, pncols, P, ldp, ccol0, cncols, Cp, ldc):
        Pv = self._v(P, ldp, pncols)
        Cv = self._v(Cp, ldc, cncols)
        d = ccol0 - pcol0
        Cv -= Pv[d:, :] @ Pv[d:d + cncols, :].T

    def panel_trsv(self, n, col0, ncols, P, ldp, y, acc, z, scratch):
        V = self._v(P, ldp, ncols)
        yv, av, zv = y.numpy(), acc.numpy(), z.numpy()
        rhs = yv[col0:col0 + ncols] - av[col0:col0 + ncols]
        zz = np.linalg.solve(np.tril(V[:ncols, :]), rhs)
        zv[col0:col0 + ncols] = zz
        av[col0 + ncols:] += V[ncols:, :] @ zz

    def panel_logdiag(self, n, col0, ncols, P, ldp, out):
        V = self._v(P, ldp, ncols)
        dg = np.diag(V[:ncols, :])
        idx = col0 + np.arange(ncols)
        out.numpy()[0] += np.sum(np.log(dg[idx < n]))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, pc, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gp_b200.block_cyclic import BlockCyclicGP
    from oracle import gp_oracle as o
    x, y = o.synth_xy(n, 5)
    bc = BlockCyclicGP(n, panel_cols=pc, backend=NumpyPanelBackend())
    info = bc.factor(x, 1.0, 1.0, 0.3)
    val = bc.lml(y)
    L = bc.gather_factor()
    q.put((rank, info, val, L))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,pc", [(2, 500, 128), (2, 700, 256), (3, 900, 128)])
def test_block_cyclic_schedule_gloo(world, n, pc):
    from oracle import gp_oracle as o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, pc, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y = o.synth_xy(n, 5)
    Lref = o.cholesky_decompose(o.gram_se(x, 1.0, 1.0, 0.09))
    vref = o.lml(x, y, 1.0, 1.0, 0.3)
    for rank, info, val, L in res:
        assert info == 0
        assert np.max(np.abs(L - Lref)) < 1e-10
        assert abs(val - vref) < 1e-10 * abs(vref)


def test_single_process_schedule():
    from gp_b200.block_cyclic import BlockCyclicGP
    from oracle import gp_oracle as o
    n = 300
    x, y = o.synth_xy(n, 5)
    bc = BlockCyclicGP(n, panel_cols=128, backend=NumpyPanelBackend())
    assert bc.factor(x, 1.2, 0.8, 0.2) == 0
    assert abs(bc.lml(y) - o.lml(x, y, 1.2, 0.8, 0.2)) < 1e-10 * abs(o.lml(x, y, 1.2, 0.8, 0.2))
