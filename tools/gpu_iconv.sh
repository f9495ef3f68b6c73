#!/bin/bash
# tcgen05 iconv1: parity tests (with a watchdog: a pipeline bug would hang), then the timing against the library path
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_iconv_gpu.py -x -q -m gpu --timeout 120 > gpurun_out/pytest_iconv.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -25 gpurun_out/pytest_iconv.log | cut -c1-300
if [ $rc -eq 0 ]; then
  timeout 300 python tools/bench_iconv.py > gpurun_out/bench_iconv.json 2> gpurun_out/bench_iconv.err; echo "bench exit $?"; cat gpurun_out/bench_iconv.json | cut -c1-2500; tail -3 gpurun_out/bench_iconv.err
fi
