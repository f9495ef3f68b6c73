#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_wgrad_gpu.py tests/test_decoder_gpu.py tests/test_trainer_gpu.py -x -q > gpurun_out/wgrad2_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/wgrad2_tests.log
timeout 900 python tools/bench_decoder.py --config 5 4 > gpurun_out/wgrad2_bench.jsonl 2> gpurun_out/wgrad2_bench.err; echo "bench rc=$?"
cut -c1-420 gpurun_out/wgrad2_bench.jsonl; tail -5 gpurun_out/wgrad2_bench.err
