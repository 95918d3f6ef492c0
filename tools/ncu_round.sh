#!/bin/bash
# ncu evidence for one round: launch list of one timed bench step + full captures of the dominant
# kernel.  Each ncu run follows a plain run of the same command that exited 0.
TAG=${1:-r01}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 336 -c 112 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tile_kernel -s 140 -c 2 \
    -o gpurun_out/prof_gemm_chol_$TAG -f $CMD > gpurun_out/ncu_chol_$TAG.log 2>&1
$CMD > gpurun_out/plain3_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tile_kernel -s 167 -c 1 \
    -o gpurun_out/prof_gemm_lauum_$TAG -f $CMD > gpurun_out/ncu_lauum_$TAG.log 2>&1
ls -la gpurun_out/
tail -2 gpurun_out/plain_$TAG.log
$CMD > gpurun_out/plain4_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gram_se_batched -c 1 \
    -o gpurun_out/prof_gram_$TAG -f $CMD > gpurun_out/ncu_gram_$TAG.log 2>&1
$CMD > gpurun_out/plain5_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tile_kernel -s 184 -c 1 \
    -o gpurun_out/prof_gemm_trtri_$TAG -f $CMD > gpurun_out/ncu_trtri_$TAG.log 2>&1
