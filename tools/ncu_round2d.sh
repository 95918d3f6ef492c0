#!/bin/bash
# --set full captures of the 128-row fused panel launch and of a fine-tile GEMM launch on the B = 1 chain
TAG=${1:-r02}
PCMD="python tools/panel_one.py 3 128 1"
$PCMD > gpurun_out/plain_fusedtile_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:panel_fused_tile_kernel -s 1 -c 1 -o gpurun_out/prof_panel_fused_tile_$TAG -f $PCMD > gpurun_out/ncu_fusedtile_$TAG.log 2>&1
LCMD="python tools/latency_one.py 2048"
GPB200_NO_GRAPH=1 $LCMD > gpurun_out/plain_fine_$TAG.log 2>&1 && \
GPB200_NO_GRAPH=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:GemmCfg<.int.4, .int.4, .int.64" -s 20 -c 1 -o gpurun_out/prof_gemm_fine_$TAG -f $LCMD > gpurun_out/ncu_fine_$TAG.log 2>&1
ls -la gpurun_out/prof_panel_fused_tile_$TAG* gpurun_out/prof_gemm_fine_$TAG*
