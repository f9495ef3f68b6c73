#!/bin/bash
# ncu session for the multi-layer kernels (default tuning): launch list + full capture
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3 --skip-e2e --skip-cpu --skip-extras --no-graph --mode multi $EXTRA"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lpg_ -s 2 -c 4 -o gpurun_out/prof_multi $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
