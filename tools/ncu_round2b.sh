#!/bin/bash
# --set full captures of the two dominant GEMM instances (template arguments are matched on the demangled name)
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tile_kernel<.*128>, .*0, .*0, .*0>" -s 20 -c 1 \
    -o gpurun_out/prof_gemm_nt_$TAG -f $CMD > gpurun_out/ncu_nt_$TAG.log 2>&1
$CMD > gpurun_out/plain3_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_tile_kernel<.*128>, .*1, .*1, .*1>" -s 1 -c 1 \
    -o gpurun_out/prof_gemm_lauum_$TAG -f $CMD > gpurun_out/ncu_lauum_$TAG.log 2>&1
ls -la gpurun_out/prof_gemm*_$TAG*; tail -3 gpurun_out/ncu_nt_$TAG.log
