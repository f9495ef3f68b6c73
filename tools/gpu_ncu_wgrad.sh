#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/prof_wgrad.py 32 16"
$CMD > gpurun_out/wgrad_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_wgrad" -s 2 -c 1 -f -o gpurun_out/prof_wgrad $CMD > gpurun_out/ncu_wgrad.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_wgrad.log
