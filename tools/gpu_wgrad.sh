#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_wgrad_gpu.py -x -q > gpurun_out/wgrad_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/wgrad_tests.log
timeout 600 python tools/bench_wgrad.py > gpurun_out/wgrad.json 2> gpurun_out/wgrad.err; echo "bench rc=$?"; tail -3 gpurun_out/wgrad.err
