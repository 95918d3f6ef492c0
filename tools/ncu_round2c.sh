#!/bin/bash
# launch list of one eager B = 1, N = 4096 evaluation (the latency path) and a --set full capture of the fused
# POTRF + TRSM launch; each after the same command has exited 0 without ncu
TAG=${1:-r02}
CMD="python tools/latency_one.py 4096"
GPB200_NO_GRAPH=1 $CMD > gpurun_out/plain_lat_$TAG.log 2>&1 && \
GPB200_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_lat4096_$TAG.csv $CMD > gpurun_out/ncu_lat_$TAG.log 2>&1
PCMD="python tools/panel_one.py 3 32 1"
$PCMD > gpurun_out/plain_fused_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:panel_fused_kernel -s 1 -c 1 -o gpurun_out/prof_panel_fused_$TAG -f $PCMD > gpurun_out/ncu_fused_$TAG.log 2>&1
ls -la gpurun_out/prof_panel_fused_$TAG* gpurun_out/launches_lat4096_$TAG.csv
