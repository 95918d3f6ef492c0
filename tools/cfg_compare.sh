#!/bin/bash
# bench the headline step under each long-k GEMM configuration (1 = one 128x128 CTA per SM, 2 = two half-tile CTAs per SM)
for c in ${CFGS:-1 2}; do
  GPB200_GEMM_CFG=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_cfg$c.json"))
print("cfg $c", round(d["value"],1), "evals/s  gemm", round(d["roofline"]["kernel_ms_per_step"],1), "ms  frac", round(d["roofline"]["frac"],4), "chol", round(d["cholesky"]["tflops_per_gpu"],2), d["roofline"]["other_kernels_ms_per_step"])
PY
done
