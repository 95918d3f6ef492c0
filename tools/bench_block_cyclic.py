"""SURVEY config 5: single large exact GP, N = 1k..32k, block-column-cyclic Cholesky over the GPUs of
one box with NCCL panel broadcasts.  Launch with torchrun (or plain python for one GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/bench_block_cyclic.py [N ...]
Prints one JSON line per N on rank 0: factorisation time (CUDA events, max over ranks), N^3/3
TFLOP/s, and the LML (checked against a one-GPU evaluation by the caller/tests)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_b200 import capi  # noqa: E402
from gp_b200.block_cyclic import BlockCyclicGP, GpuPanelBackend  # noqa: E402


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
    be = GpuPanelBackend(h, dev)
    out = []
    for n in sizes:
        rng = np.random.default_rng(5)
        x = np.sort(rng.uniform(0, 0.05 * n, n))
        y = np.sin(x) + 0.5 * np.sin(3.1 * x) + 0.3 * rng.standard_normal(n)
        bc = BlockCyclicGP(n, panel_cols=pc, backend=be)
        bc.factor(x, 1.0, 1.0, 0.3)  # warm-up (task lists, NCCL channels)
        best = 1e30
        for _ in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            info = bc.factor(x, 1.0, 1.0, 0.3)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = min(best, float(t.item()))
        t0 = time.perf_counter()
        val = bc.lml(y)
        torch.cuda.synchronize()
        lml_ms = (time.perf_counter() - t0) * 1e3
        rec = {"n": n, "gpus": world, "panel_cols": pc, "factor_ms": round(best, 3),
               "chol_tflops": round(n ** 3 / 3.0 / best * 1e-9, 2), "lml_ms": round(lml_ms, 2), "lml": val, "info": info}
        out.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        del bc
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/block_cyclic_%dgpu.json" % world, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
