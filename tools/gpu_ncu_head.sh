#!/bin/bash
# ncu capture of the fused head kernels at r = $R (default 8), f32, both encoders
mkdir -p gpurun_out
R=${R:-8}
CMD="python tools/sweep_head.py --only-r $R"
$CMD > gpurun_out/head_plain.json 2> gpurun_out/head_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:"head_lpg_bwd" -s 2 -c 2 -f -o gpurun_out/prof_head_r$R $CMD > gpurun_out/ncu_head.log 2>&1
echo "ncu head exit $?"; tail -2 gpurun_out/ncu_head.log
