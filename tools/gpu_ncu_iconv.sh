#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/prof_iconv.py 32"
$CMD > gpurun_out/iconv_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"iconv1_fwd" -s 2 -c 1 -f -o gpurun_out/prof_iconv $CMD > gpurun_out/ncu_iconv.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_iconv.log
