"""Block-column-cyclic multi-GPU Cholesky / LML of ONE large exact GP.

The reference has nothing like this (its answer to large N is approximation); the
semantics are those of models/fit_hyperparameters.stan:19-31 at a size where one GPU is too slow.

Layout: the padded matrix (np = ceil(n/128)*128) is cut into panels of `panel_cols` columns; panel p
belongs to rank p % world and is stored compactly (rows [p*panel_cols, np) only).  Every rank builds
its own panels of the Gram matrix from the replicated x (no K scatter).

Algorithm (right-looking, look-ahead 1):
    for p in panels:
        owner(p) has factored panel p (diagonal block + rows below) -> NCCL broadcast of the panel
        owner(p+1) first applies panel p to panel p+1, factors it and starts ITS broadcast
        everybody applies panel p to the rest of its own panels (DMMA GEMM, K = panel_cols)
so the broadcast of panel p+1 travels over NVLink while the trailing update of panel p runs.
The only collectives are those panel broadcasts, the small replicated vectors of the forward
substitution and one all-reduce of (log-det, info) -- torch.distributed (NCCL on GPUs).

The per-rank compute is behind a tiny backend interface so that the schedule (ownership, ordering,
buffers, collectives) can be exercised with world_size-2 gloo processes on CPU
(tests/test_block_cyclic.py); on GPUs the backend is the C ABI of include/gpb200.h (gpb200_mg_*).
"""
from __future__ import annotations

import ctypes as C
import math

import os

import numpy as np

TILE = 128


def padded(n: int) -> int:
    return (n + TILE - 1) // TILE * TILE


class GpuPanelBackend:
    """Per-rank compute through libgpb200.so on torch CUDA tensors (device-pointer ABI)."""

    def __init__(self, handle, device):
        import torch
        self.torch = torch
        self.h = handle
        self.device = device
        self.lib = handle.lib
        handle.set_stream(torch.cuda.current_stream(device).cuda_stream)

    def _chk(self, rc, where):
        self.h._check(rc, where, allow_info=False)

    def empty(self, rows, cols):
        return self.torch.empty((cols, rows), dtype=self.torch.float64, device=self.device)  # column-major rows x cols

    def vector(self, n, zero=True):
        f = self.torch.zeros if zero else self.torch.empty
        return f(n, dtype=self.torch.float64, device=self.device)

    def from_host(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    def info_scalar(self):
        return self.torch.zeros(1, dtype=self.torch.int32, device=self.device)

    def gram_panel(self, n, x, alpha, rho, diag_add, col0, ncols, P, ldp):
        self._chk(self.lib.gpb200_mg_gram_panel(self.h._h, n, x.data_ptr(), alpha, rho, diag_add, col0, ncols,
                                                P.data_ptr(), ldp), "mg_gram_panel")

    def panel_factor(self, n, col0, ncols, P, ldp, info):
        self._chk(self.lib.gpb200_mg_panel_factor(self.h._h, n, col0, ncols, P.data_ptr(), ldp, info.data_ptr()),
                  "mg_panel_factor")

    def panel_factor_col(self, n, col0, ncols, P, ldp, jl, info):
        self._chk(self.lib.gpb200_mg_panel_factor_col(self.h._h, n, col0, ncols, P.data_ptr(), ldp, jl, info.data_ptr()),
                  "mg_panel_factor_col")

    def use_stream(self, stream, ordered=False):
        """The block-cyclic schedule orders its streams with events of its own; an ordering edge on every switch would
        serialise the panel chain with the trailing updates."""
        self.h.set_stream(stream.cuda_stream, ordered=ordered)

    def panel_update(self, n, pcol0, pncols, P, ldp, ccol0, cncols, Cp, ldc, jl0=0, jl1=None):
        jl1 = cncols // TILE if jl1 is None else jl1
        self._chk(self.lib.gpb200_mg_panel_update_cols(self.h._h, n, pcol0, pncols, P.data_ptr(), ldp, ccol0, cncols,
                                                       Cp.data_ptr(), ldc, jl0, jl1), "mg_panel_update")

    def panel_trsv(self, n, col0, ncols, P, ldp, y, acc, z, scratch):
        self._chk(self.lib.gpb200_mg_panel_trsv(self.h._h, n, col0, ncols, P.data_ptr(), ldp, y.data_ptr(),
                                                acc.data_ptr(), z.data_ptr(), scratch.data_ptr()), "mg_panel_trsv")

    def panel_logdiag(self, n, col0, ncols, P, ldp, out):
        self._chk(self.lib.gpb200_mg_panel_logdiag(self.h._h, n, col0, ncols, P.data_ptr(), ldp, out.data_ptr()),
                  "mg_panel_logdiag")

    def to_host(self, t):
        return t.cpu().numpy()

    # -- collectives enqueued from C (ncclBroadcast / ncclAllReduce on the handle's streams) -------------
    def init_comm(self, group=None):
        """Creates the handle's NCCL communicator: rank 0 draws the unique id in C (gpb200_mg_comm_id), the
        128 bytes travel once through torch.distributed, every rank calls gpb200_mg_comm_init."""
        import torch.distributed as dist
        torch = self.torch
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        ident = torch.zeros(128, dtype=torch.uint8)
        if world > 1:
            if rank == 0:
                buf = (C.c_char * 128)()
                self._chk(self.lib.gpb200_mg_comm_id(self.h._h, buf), "mg_comm_id")
                ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            ident = ident.to(self.device)
            dist.broadcast(ident, src=0, group=group)
            ident = ident.cpu()
        raw = (C.c_char * 128).from_buffer_copy(bytes(ident.numpy().tobytes()))
        self._chk(self.lib.gpb200_mg_comm_init(self.h._h, raw, rank, world), "mg_comm_init")
        self.native_comm = True
        return self

    def close_comm(self):
        if getattr(self, "native_comm", False):
            self.lib.gpb200_mg_comm_destroy(self.h._h)
            self.native_comm = False

    def bcast(self, buf, count, root):
        t = C.c_longlong(0)
        self._chk(self.lib.gpb200_mg_bcast(self.h._h, buf.data_ptr(), int(count), int(root), C.byref(t)), "mg_bcast")
        return t.value

    def wait(self, ticket):
        self._chk(self.lib.gpb200_mg_wait(self.h._h, int(ticket)), "mg_wait")

    def allreduce(self, t, op="sum"):
        is_int = t.dtype == self.torch.int32
        self._chk(self.lib.gpb200_mg_allreduce(self.h._h, t.data_ptr(), t.numel(), 1 if op == "max" else 0, int(is_int)),
                  "mg_allreduce")

    # -- distributed gradient building blocks ---------------------------------------------------------------
    def my_columns(self, n, pc, rank, world):
        return int(self.lib.gpb200_mg_my_columns(n, pc, rank, world))

    def panel_to_square(self, n, col0, ncols, P, ldp, Lsq):
        self._chk(self.lib.gpb200_mg_panel_to_square(self.h._h, n, col0, ncols, P.data_ptr(), ldp, Lsq.data_ptr()),
                  "mg_panel_to_square")

    def inverse_rows(self, n, pc, rank, world, Lsq, Xp, S, Wd):
        self._chk(self.lib.gpb200_mg_inverse_rows(self.h._h, n, pc, rank, world, Lsq.data_ptr(), Xp.data_ptr(), S.data_ptr(),
                                                  Wd.data_ptr()), "mg_inverse_rows")

    def solve_partials(self, n, pc, rank, world, Lsq, Xp, ypad, z_mine, a_part, sums2, part):
        self._chk(self.lib.gpb200_mg_solve_partials(self.h._h, n, pc, rank, world, Lsq.data_ptr(), Xp.data_ptr(), ypad.data_ptr(),
                                                    z_mine.data_ptr(), a_part.data_ptr(), sums2.data_ptr(), part.data_ptr()),
                  "mg_solve_partials")

    def quadform_partials(self, n, rank, world, x, a, theta3, part, sums2):
        self._chk(self.lib.gpb200_mg_quadform_partials(self.h._h, n, rank, world, x.data_ptr(), a.data_ptr(), theta3.data_ptr(),
                                                       part.data_ptr(), sums2.data_ptr()), "mg_quadform_partials")

    def trace_partials(self, n, pc, rank, world, Xp, x, avec, theta3, partial, sums3):
        self._chk(self.lib.gpb200_mg_trace_partials(self.h._h, n, pc, rank, world, Xp.data_ptr(), x.data_ptr(), avec.data_ptr(),
                                                    theta3.data_ptr(), partial.data_ptr(), sums3.data_ptr()), "mg_trace_partials")


class BlockCyclicGP:
    """Distributed factorisation and LML of one exact GP with a squared-exponential kernel."""

    def __init__(self, n, panel_cols=512, backend=None, group=None, keep_all=False, split_update=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n = int(n)
        self.np_ = padded(self.n)
        self.pc = min(int(panel_cols), self.np_)
        if self.pc % TILE:
            raise ValueError("panel_cols must be a multiple of 128")
        self.npanels = (self.np_ + self.pc - 1) // self.pc
        self.be = backend
        self.panels = {}
        self.info = None
        # keep_all: every rank keeps every broadcast panel (the gradient needs the whole factor on every rank; the
        # panels arrive anyway).  Otherwise two receive buffers are recycled.
        self.keep_all = bool(keep_all)
        # split_update (native schedule): the owner of the next panel applies the arriving panel to its first tile column on
        # the panel chain and to the others on a side stream
        if split_update is None:
            split_update = os.environ.get("GPB200_MG_SPLIT_UPDATE", "1") != "0"
        self.split_update = bool(split_update)
        self.received = {}
        self.native = bool(getattr(backend, "native_comm", False))
        self.x_dev = None
        self.theta = None

    # -- geometry -----------------------------------------------------------------------------
    def owner(self, p):
        return p % self.world

    def col0(self, p):
        return p * self.pc

    def ncols(self, p):
        return min(self.pc, self.np_ - p * self.pc)

    def ld(self, p):
        return self.np_ - p * self.pc

    def my_panels(self):
        return [p for p in range(self.npanels) if self.owner(p) == self.rank]

    def _bcast(self, tensor, src, async_op=False):
        if self.world == 1:
            return None
        if self.native:   # enqueued from C on the handle's communication stream; the "work" is a ticket
            ticket = self.be.bcast(tensor, tensor.numel(), src)
            if async_op:
                return ticket
            self.be.wait(ticket)
            return None
        return self.dist.broadcast(tensor, src=src, group=self.group, async_op=async_op)

    def _wait(self, work):
        if work is None:
            return
        if self.native:
            self.be.wait(work)
        else:
            work.wait()

    # -- factorisation ------------------------------------------------------------------------
    def factor(self, x, alpha, rho, sigma, jitter=0.0):
        """K = cov_exp_quad(x, alpha, rho) + (sigma^2 + jitter) I  ->  L, distributed.  Returns LAPACK info."""
        be, n = self.be, self.n
        dx = be.from_host(x)
        self.x_dev = dx
        self.theta = (float(alpha), float(rho), float(sigma), float(jitter))
        self.info = be.info_scalar()
        self.panels = {}
        self.received = {}
        self._square = None
        if self.native:
            return self._factor_native(dx, alpha, rho, sigma, jitter)
        for p in self.my_panels():
            P = be.empty(self.ld(p), self.ncols(p))
            be.gram_panel(n, dx, float(alpha), float(rho), float(sigma) ** 2 + float(jitter), self.col0(p),
                          self.ncols(p), P, self.ld(p))
            self.panels[p] = P
        # receive buffers for the panel being applied and the one in flight (look-ahead 1)
        recv = [None, None]
        pending = {}

        def start_bcast(p):
            own = self.owner(p)
            if own == self.rank:
                buf = self.panels[p]
            elif self.keep_all:
                buf = be.empty(self.ld(p), self.ncols(p))
                self.received[p] = buf
            else:
                slot = p % 2
                need = self.ld(p) * self.ncols(p)
                if recv[slot] is None or recv[slot].numel() < need:
                    recv[slot] = be.empty(self.ld(0), self.pc)
                buf = recv[slot].reshape(-1)[:need].reshape(self.ncols(p), self.ld(p))
            work = self._bcast(buf, own, async_op=True)
            pending[p] = (buf, work)

        if self.owner(0) == self.rank:
            be.panel_factor(n, self.col0(0), self.ncols(0), self.panels[0], self.ld(0), self.info)
        start_bcast(0)
        for p in range(self.npanels):
            buf, work = pending.pop(p)
            self._wait(work)
            nxt = p + 1
            if nxt < self.npanels:
                if self.owner(nxt) == self.rank:
                    be.panel_update(n, self.col0(p), self.ncols(p), buf, self.ld(p), self.col0(nxt), self.ncols(nxt),
                                    self.panels[nxt], self.ld(nxt))
                    be.panel_factor(n, self.col0(nxt), self.ncols(nxt), self.panels[nxt], self.ld(nxt), self.info)
                start_bcast(nxt)
            for q in self.my_panels():
                if q > nxt:
                    be.panel_update(n, self.col0(p), self.ncols(p), buf, self.ld(p), self.col0(q), self.ncols(q),
                                    self.panels[q], self.ld(q))
        info = self.info.clone()
        if self.world > 1:
            # first failing pivot over all ranks: max over ranks of (info>0 ? BIG - info : 0) keeps the smallest
            big = 1 << 30
            enc = (info > 0).to(info.dtype) * (big - info)
            if self.native:
                self.be.allreduce(enc, "max")
            else:
                self.dist.all_reduce(enc, op=self.dist.ReduceOp.MAX, group=self.group)
            info = (enc > 0).to(info.dtype) * (big - enc)
        return int(be.to_host(info)[0])

    def _factor_native(self, dx, alpha, rho, sigma, jitter):
        """The same right-looking schedule with everything enqueued from C and nothing waiting on the host:
          * the panel chain -- apply panel p to panel p+1, factor it tile column by tile column, hand each finished
            tile column to ncclBroadcast -- runs on a HIGH-PRIORITY stream, so it overtakes the rank's own backlog of
            trailing updates (on the main stream) instead of queueing behind it;
          * a tile column travels while the next one is being factored (four broadcasts of ldp x 128 per 512-panel);
          * every rank waits for a panel only where it first reads it."""
        be, n = self.be, self.n
        torch = be.torch
        main = torch.cuda.current_stream(be.device)
        if getattr(self, "_pstream", None) is None:
            self._pstream = torch.cuda.Stream(device=be.device, priority=-2)
            self._ustream = torch.cuda.Stream(device=be.device, priority=-1)
        ps, us = self._pstream, self._ustream
        split = self.split_update
        be.use_stream(main)
        for p in self.my_panels():
            P = be.empty(self.ld(p), self.ncols(p))
            be.gram_panel(n, dx, float(alpha), float(rho), float(sigma) ** 2 + float(jitter), self.col0(p), self.ncols(p), P,
                          self.ld(p))
            self.panels[p] = P
        recv = [None, None]
        last_ticket = [None]
        square = None
        self._square = None
        if self.keep_all:
            square = be.empty(self.np_, self.np_)
        last_upd = {}          # my panel q -> event after the latest trailing update written into it (main stream)
        built = torch.cuda.Event()
        built.record(main)

        def buffer_of(p):
            if self.owner(p) == self.rank:
                return self.panels[p]
            if self.keep_all:
                self.received[p] = be.empty(self.ld(p), self.ncols(p))
                return self.received[p]
            slot = p % 2
            need = self.ld(p) * self.ncols(p)
            if recv[slot] is None or recv[slot].numel() < need:
                recv[slot] = be.empty(self.ld(0), self.pc)
            return recv[slot].reshape(-1)[:need].reshape(self.ncols(p), self.ld(p))

        def factor_and_bcast(p, prev_buf, prev_tickets):
            """panel stream: (apply panel p-1) -> factor tile columns -> broadcast each; returns (buffer, tickets)"""
            own = self.owner(p)
            buf = buffer_of(p)
            ntc = self.ncols(p) // TILE
            chunk = self.ld(p) * TILE
            flat = buf.reshape(-1)
            tickets = []
            col_ready = {}
            be.use_stream(ps)
            if own == self.rank:
                ps.wait_event(last_upd.get(p, built))
                if prev_buf is not None:
                    for t in prev_tickets:
                        be.wait(t)
                    upd = (n, self.col0(p - 1), self.ncols(p - 1), prev_buf, self.ld(p - 1), self.col0(p), self.ncols(p), buf,
                           self.ld(p))
                    if split and ntc > 1:
                        # only tile column 0 has to be updated before the chain can go on; the other tile columns are
                        # updated beside it (the POTRF / TRSM kernels of the chain leave most SMs idle)
                        ready = torch.cuda.Event()
                        ready.record(ps)
                        us.wait_event(ready)
                        be.use_stream(us)
                        for jl in range(1, ntc):
                            be.panel_update(*upd, jl0=jl, jl1=jl + 1)
                            col_ready[jl] = torch.cuda.Event()
                            col_ready[jl].record(us)
                        be.use_stream(ps)
                        be.panel_update(*upd, jl0=0, jl1=1)
                    else:
                        be.panel_update(*upd)
            elif not self.keep_all:
                # a recycled receive buffer may still be read by trailing updates of panel p-2 on the main stream
                ev = torch.cuda.Event(); ev.record(main); ps.wait_event(ev)
            for jl in range(ntc):
                if jl in col_ready:
                    ps.wait_event(col_ready[jl])
                if own == self.rank:
                    be.panel_factor_col(n, self.col0(p), self.ncols(p), buf, self.ld(p), jl, self.info)
                tickets.append(be.bcast(flat[jl * chunk:(jl + 1) * chunk], chunk, own))
            last_ticket[0] = tickets[-1]
            be.use_stream(main)
            return buf, tickets

        cur_buf, cur_t = factor_and_bcast(0, None, None)
        for p in range(self.npanels):
            nxt = p + 1
            nxt_buf = nxt_t = None
            if nxt < self.npanels:
                nxt_buf, nxt_t = factor_and_bcast(nxt, cur_buf, cur_t)
            be.use_stream(main)
            mine = [q for q in self.my_panels() if q > nxt]
            if mine or square is not None:
                for t in cur_t:
                    be.wait(t)
                for q in mine:
                    be.panel_update(n, self.col0(p), self.ncols(p), cur_buf, self.ld(p), self.col0(q), self.ncols(q),
                                    self.panels[q], self.ld(q))
                    ev = torch.cuda.Event()
                    ev.record(main)
                    last_upd[q] = ev
                if square is not None:
                    # the gradient wants the factor as one square: file the panel now, beside the chain, instead of in
                    # 2 N^2 words of copies at the start of lml_grad
                    be.panel_to_square(n, self.col0(p), self.ncols(p), cur_buf, self.ld(p), square)
            cur_buf, cur_t = nxt_buf, nxt_t
        done = torch.cuda.Event()
        done.record(ps)
        main.wait_event(done)
        # every broadcast has to have landed before anyone reads the received panels (the communication stream runs
        # them in order: waiting for the last ticket covers all)
        be.use_stream(main)
        if last_ticket[0] is not None:
            be.wait(last_ticket[0])
        self._square = square
        info = self.info.clone()
        if self.world > 1:
            big = 1 << 30
            enc = (info > 0).to(info.dtype) * (big - info)
            be.allreduce(enc, "max")
            info = (enc > 0).to(info.dtype) * (big - enc)
        return int(be.to_host(info)[0])

    # -- likelihood AND gradient ---------------------------------------------------------------------
    def lml_grad(self, y):
        """MVN(y | 0, K) log density and its gradient in (alpha, rho, sigma) (models/fit_hyperparameters.stan:19-31 and
        its reverse sweep) from the distributed factor.  Needs keep_all=True (every rank holds every panel).  Exact:
        rank r inverts the rows of L^-1 that belong to its panels and contracts K^-1's contribution of those rows with
        dK/dtheta in the fused trace epilogue; two small all-reduces combine the ranks.  No stochastic estimator."""
        if not self.keep_all:
            raise ValueError("lml_grad needs BlockCyclicGP(..., keep_all=True)")
        be, n, np_, pc = self.be, self.n, self.np_, self.pc
        torch = be.torch
        alpha, rho, sigma, _ = self.theta
        Lsq = getattr(self, "_square", None)
        if Lsq is None:   # host-driven schedule: assemble the square now
            Lsq = be.empty(np_, np_)
            for p in range(self.npanels):
                P = self.panels[p] if self.owner(p) == self.rank else self.received[p]
                be.panel_to_square(n, self.col0(p), self.ncols(p), P, self.ld(p), Lsq)
        nmine = be.my_columns(n, pc, self.rank, self.world)
        Xp = be.empty(np_, max(nmine, 1))
        S = be.empty(pc, max(nmine, 1))
        Wd = be.empty(pc * pc, 2 * self.npanels)
        ypad = be.from_host(np.concatenate([np.asarray(y, dtype=np.float64), np.zeros(np_ - n)]))
        z_mine = be.vector(max(nmine, 1))
        a = be.vector(np_)
        sums2 = be.vector(2)
        part = be.vector(8 * np_, zero=False)
        be.inverse_rows(n, pc, self.rank, self.world, Lsq, Xp, S, Wd)
        be.solve_partials(n, pc, self.rank, self.world, Lsq, Xp, ypad, z_mine, a, sums2, part)
        del S, Wd
        Lsq = None
        qf = sums2[0:1].clone()
        if self.world > 1:
            be.allreduce(a)
            be.allreduce(qf)
        nt = np_ // TILE
        partial = be.vector(16 * nt * (nt + 1) // 2, zero=False)
        sums5 = be.vector(5)    # (-sum G e, -sum G e d^2, tr G) of this rank's rows of L^-1, then (a^T E a, a^T (E o D2) a) of its slice
        theta3 = be.from_host(np.array([alpha, rho, sigma]))
        be.trace_partials(n, pc, self.rank, self.world, Xp, self.x_dev, be.vector(np_), theta3, partial, sums5[0:3])
        be.quadform_partials(n, self.rank, self.world, self.x_dev, a, theta3, partial, sums5[3:5])
        if self.world > 1:
            be.allreduce(sums5)
        aa = float((a[:n] * a[:n]).sum().item())
        s5 = be.to_host(sums5)
        s_se, s_d2, s_tr = float(s5[0] + s5[3]), float(s5[1] + s5[4]), float(s5[2])
        logdet = float(be.to_host(sums2)[1])
        lml = -0.5 * n * math.log(2.0 * math.pi) - logdet - 0.5 * float(qf.item())
        grad = np.array([alpha * s_se, 0.5 * alpha * alpha * s_d2 / rho ** 3, sigma * (aa - s_tr)])
        return lml, grad

    # -- likelihood ---------------------------------------------------------------------------
    def lml(self, y):
        """MVN(y | 0, K) log density with constants (models/fit_hyperparameters.stan:31) from the
        distributed factor: forward substitution panel by panel (owner solves, replicated vectors are
        re-broadcast), log-det all-reduced."""
        be, n = self.be, self.n
        dy = be.from_host(np.concatenate([np.asarray(y, dtype=np.float64), np.zeros(self.np_ - n)]))
        acc = be.vector(self.np_)
        z = be.vector(self.np_)
        logdet = be.vector(1)
        scratch = None
        for p in range(self.npanels):
            own = self.owner(p)
            c0, nc = self.col0(p), self.ncols(p)
            if own == self.rank:
                if scratch is None:
                    scratch = be.empty(self.ld(0), TILE)
                be.panel_trsv(n, c0, nc, self.panels[p], self.ld(p), dy, acc, z, scratch)
                be.panel_logdiag(n, c0, nc, self.panels[p], self.ld(p), logdet)
            if self.world > 1:
                self._bcast(z[c0:c0 + nc], own)
                if c0 + nc < self.np_:
                    self._bcast(acc[c0 + nc:], own)
        if self.world > 1:
            self.dist.all_reduce(logdet, op=self.dist.ReduceOp.SUM, group=self.group)
        zh = be.to_host(z)[:n]
        ld = float(be.to_host(logdet)[0])
        return -0.5 * n * math.log(2.0 * math.pi) - ld - 0.5 * float(zh @ zh)

    def gather_factor(self):
        """Dense lower factor on the host of every rank (tests only; O(N^2) traffic)."""
        L = np.zeros((self.np_, self.np_))
        for p in range(self.npanels):
            own = self.owner(p)
            c0, nc, ld = self.col0(p), self.ncols(p), self.ld(p)
            if own == self.rank:
                buf = self.panels[p]
            else:
                buf = self.be.empty(ld, nc)
            self._bcast(buf, own)
            blk = self.be.to_host(buf).reshape(nc, ld).T   # rows [c0, np) x nc
            L[c0:, c0:c0 + nc] = blk
        return np.tril(L)[:self.n, :self.n]
