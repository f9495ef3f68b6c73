#!/bin/bash
# First GPU session: parity tests, bench (f32, bf16), tuning sweep, then (only if the tests passed)
# the ncu launch list and one full capture of the multi-layer kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
rc=$?
echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then
  python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_all.log 2>&1
  tail -40 gpurun_out/pytest_gpu_all.log
fi
timeout 600 python bench.py > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; echo "bench f32 exit $?"; cat gpurun_out/bench_f32.json | head -c 3000
timeout 300 python bench.py --dtype bf16 --skip-cpu > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit $?"
timeout 600 python tools/sweep.py > gpurun_out/sweep.json 2> gpurun_out/sweep.err; echo "sweep exit $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
if [ $rc -eq 0 ]; then
  CMD="python bench.py --steps 5 --warmup 3 --skip-e2e --skip-cpu --skip-extras --skip-decoder --no-graph --mode multi"
  $CMD > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches exit $?"
  $CMD > gpurun_out/plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:lpg_ -s 2 -c 4 -o gpurun_out/prof_multi $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
ls -la gpurun_out
