"""C4 at its stated per-GPU scale (256 groups over 8 GPUs = 32 groups of N = 1024 per GPU) under tuning knobs."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from gp_b200 import capi
    from tools.bench_configs import run
    dev = torch.device("cuda", 0); h = capi.Handle(0); st = torch.cuda.current_stream(dev)
    h.set_stream(st.cuda_stream); h.set_pointer_mode(True)
    out = [run(h, st, dev, 1024, 32, True, reps=5), run(h, st, dev, 1024, 16, True, reps=5), run(h, st, dev, 2048, 8, True, reps=3)]
    print(json.dumps([{k: r[k] for k in ("n", "B", "ms_per_batch", "tflops")} for r in out]))
    sys.exit(0)
for label, env in (("default", {}), ("lookahead_b64", {"GPB200_LOOKAHEAD_MAXB": "64"}), ("quarter", {"GPB200_GEMM_CFG": "3"}),
                   ("lookahead_b64_quarter", {"GPB200_LOOKAHEAD_MAXB": "64", "GPB200_GEMM_CFG": "3"}),
                   ("lookahead_b64_qw4", {"GPB200_LOOKAHEAD_MAXB": "64", "GPB200_QUARTER_WAVES": "4"})):
    e = dict(os.environ); e.update(env); e["PYTHONPATH"] = "."
    r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
    print(label, r.stdout.strip().splitlines()[-1] if r.returncode == 0 else r.stderr[-800:], flush=True)
