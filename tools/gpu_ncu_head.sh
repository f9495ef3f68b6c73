#!/bin/bash
# ncu capture of the fused head backward at r = 8, f32 (densenet161 C = 128 and resnet50 C = 64)
mkdir -p gpurun_out
CMD="python tools/sweep_head.py --only-r 8 --f32-only"
$CMD > gpurun_out/head_plain.json 2> gpurun_out/head_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:"head_lpg_bwd8" -s 12 -c 1 -f -o gpurun_out/prof_head_bwd8 $CMD > gpurun_out/ncu_head.log 2>&1
echo "ncu head exit $?"; tail -2 gpurun_out/ncu_head.log
