#!/bin/bash
# ncu session for the multi-layer kernels (default tuning): parity tests, plain bench lines, launch list,
# then one full capture per dtype (each only after the same command exited 0 without ncu).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -3 gpurun_out/pytest_gpu.log
[ $rc -ne 0 ] && exit $rc
timeout 600 python bench.py > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; echo "bench f32 exit $?"
timeout 300 python bench.py --dtype bf16 --skip-cpu > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit $?"
CMD="python bench.py --steps 5 --warmup 3 --skip-e2e --skip-cpu --skip-extras --skip-decoder --no-graph --mode multi"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lpg_ -s 2 -c 4 -f -o gpurun_out/prof_multi $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full f32 exit $?"
$CMD --dtype bf16 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lpg_ -s 2 -c 4 -f -o gpurun_out/prof_multi_bf16 $CMD --dtype bf16 > gpurun_out/ncu_full_bf16.log 2>&1
echo "ncu full bf16 exit $?"
cat gpurun_out/bench_f32.json | head -c 2500
