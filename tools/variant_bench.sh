#!/bin/bash
# compares experimental builds of libgpb200 (gp_b200/lib/libgpb200_<tag>.so) on the headline bench
for tag in "" "$@"; do
  lib=gp_b200/lib/libgpb200${tag:+_$tag}.so
  [ -f "$lib" ] || continue
  GPB200_LIB=$PWD/$lib python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lml_grad_batched_shared or potrf_matches" 2>&1 | tail -1
  GPB200_LIB=$PWD/$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant=${tag:-base}', 'evals/s=%.1f'%d['value'], 'ms/step=%.1f'%d['ms_per_step'], 'gemm_ms=%.1f'%d['roofline']['kernel_ms_per_step'], 'gemm_TF=%.2f'%d['roofline']['achieved'])"
done
