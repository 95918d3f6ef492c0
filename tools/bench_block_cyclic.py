"""SURVEY config 5: ONE large exact GP, N = 4k..32k, block-column-cyclic Cholesky over the GPUs of one box with the
panel broadcasts enqueued from C (ncclBroadcast through gpb200_mg_bcast), then the exact distributed gradient.
Launch with torchrun (or plain python for one GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/bench_block_cyclic.py [N ...] [--panel=512] [--combos=n:panel_cols:split,...]
One JSON line per N on rank 0: factorisation and LML+gradient times (CUDA events, max over ranks), TFLOP/s on N^3/3 and N^3,
fraction of P x the measured FP64 peak, and the relative difference from the single-GPU path of the same library (rank 0)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_b200 import capi  # noqa: E402
from gp_b200.block_cyclic import BlockCyclicGP, GpuPanelBackend  # noqa: E402

PEAK = 37.03


def run_one(n, pc, be, h, dev, world, rank, reps=3, check_single=True, split_update=None):
    rng = np.random.default_rng(5)
    x = np.sort(rng.uniform(0, 0.05 * n, n))
    y = np.sin(x) + 0.5 * np.sin(3.1 * x) + 0.3 * rng.standard_normal(n)
    theta = (1.0, 1.0, 0.3)
    bc = BlockCyclicGP(n, panel_cols=pc, backend=be, keep_all=True, split_update=split_update)
    bc.factor(x, *theta)          # warm-up (task lists, NCCL channels)
    bc.lml_grad(y)

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    best_f, best_g = 1e30, 1e30
    for _ in range(reps):
        tf, info = timed(lambda: bc.factor(x, *theta))
        tg, (val, grad) = timed(lambda: bc.lml_grad(y))
        best_f, best_g = min(best_f, tf), min(best_g, tg)
    rec = {"n": n, "gpus": world, "panel_cols": pc, "factor_ms": round(best_f, 3), "grad_ms": round(best_g, 3),
           "ms": round(best_f + best_g, 3), "chol_tflops": round(n ** 3 / 3.0 / best_f * 1e-9, 2),
           "tflops": round(float(n) ** 3 / (best_f + best_g) * 1e-9, 2),
           "frac": round(float(n) ** 3 / (best_f + best_g) * 1e-9 / (world * PEAK), 4), "lml": val, "grad": [float(g) for g in grad],
           "grad_full": [float(g) for g in grad], "info": info, "split_update": bool(bc.split_update)}
    del bc
    torch.cuda.empty_cache()
    if check_single and rank == 0 and n <= 32768:
        h.set_pointer_mode(False)
        v1, g1 = h.lml_grad(x, y, theta)
        rec["relerr_lml_vs_single_gpu"] = float(abs(val - v1) / abs(v1))
        rec["relerr_grad_vs_single_gpu"] = float(np.max(np.abs(np.asarray(grad) - g1)) / np.max(np.abs(g1)))
    return rec, (x, y, theta)


def main():
    sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [4096, 8192, 16384, 32768]
    pc = 512
    for a in sys.argv[1:]:
        if a.startswith("--panel="):
            pc = int(a.split("=")[1])
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = capi.Handle(local)
    be = GpuPanelBackend(h, dev).init_comm()
    out = []
    combos = [(n, pc, None) for n in sizes]
    for a in sys.argv[1:]:
        if a.startswith("--combos="):       # n:panel_cols:split_update, ... -- a sweep inside one process group
            combos = [tuple(int(v) for v in c.split(":")) for c in a.split("=")[1].split(",")]
    for n, pc, sp in combos:
        rec, _ = run_one(n, pc, be, h, dev, world, rank, split_update=None if sp is None else bool(sp),
                         check_single=sp is None)
        rec.pop("grad_full", None)
        out.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/block_cyclic_%dgpu.json" % world, "w"), indent=1)
    be.close_comm()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
