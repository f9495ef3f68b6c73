#!/bin/bash
# ncu capture of the decoder-tail kernels (after the same command ran clean without ncu)
mkdir -p gpurun_out
CMD="python tools/bench_tail.py --steps 8 --warmup 2 --no-graph --skip-cpu --skip-literal $EXTRA"
$CMD > gpurun_out/tail_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"silog|eval_metrics" -s 12 -c 6 -f -o gpurun_out/prof_tail $CMD > gpurun_out/ncu_tail.log 2>&1
echo "ncu tail exit $?"; tail -3 gpurun_out/ncu_tail.log
