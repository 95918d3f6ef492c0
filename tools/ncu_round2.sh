#!/bin/bash
# ncu evidence for round 2 (one GPU).  Every ncu run follows a plain run of the same command that exited 0.
#   launches_r02.csv        every launch of one timed bench step with its device time (cold-cache, serialised: compare shares)
#   prof_gemm_nt_r02        --set full of the dominant kernel, the half-tile NT instance (Cholesky update)
#   prof_gemm_lauum_r02     --set full of the TN + fused-trace instance
#   prof_potrf_pw_r02, prof_small_r02   the round-2 panel-warp POTRF tile and the one-CTA whole-evaluation kernel
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 408 -c 140 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:gemm_tile_kernel<gpb::GemmCfg<4, 2, 64, 3, 2, 128>, 0, 0, 0>" -s 50 -c 1 \
    -o gpurun_out/prof_gemm_nt_$TAG -f $CMD > gpurun_out/ncu_nt_$TAG.log 2>&1
$CMD > gpurun_out/plain3_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:gemm_tile_kernel<gpb::GemmCfg<4, 2, 64, 3, 2, 128>, 1, 1, 1>" -s 1 -c 1 \
    -o gpurun_out/prof_gemm_lauum_$TAG -f $CMD > gpurun_out/ncu_lauum_$TAG.log 2>&1
$CMD > gpurun_out/plain4_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:potrf_tile_pw_kernel -s 40 -c 1 \
    -o gpurun_out/prof_potrf_pw_$TAG -f $CMD > gpurun_out/ncu_potrf_$TAG.log 2>&1
$CMD > gpurun_out/plain5_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lml_small_kernel -s 8 -c 1 \
    -o gpurun_out/prof_small_$TAG -f $CMD > gpurun_out/ncu_small_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG*
