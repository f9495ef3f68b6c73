#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_daspp_train_gpu.py tests/test_bnglue_gpu.py -x -q > gpurun_out/daspp_tests.log 2>&1; echo "daspp rc=$?"
tail -40 gpurun_out/daspp_tests.log
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_trainer_gpu.py -x -q > gpurun_out/daspp_decoder.log 2>&1; echo "decoder rc=$?"
tail -15 gpurun_out/daspp_decoder.log
timeout 900 python tools/bench_decoder.py --config 5 4 > gpurun_out/daspp_bench.jsonl 2> gpurun_out/daspp_bench.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/daspp_bench.jsonl; tail -5 gpurun_out/daspp_bench.err
