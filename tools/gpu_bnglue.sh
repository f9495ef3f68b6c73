#!/bin/bash
# BN training glue: parity tests, decoder fixtures, then the decoder training bench (configs 5 and 4)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bnglue_gpu.py -x -q > gpurun_out/bnglue_tests.log 2>&1; echo "bnglue rc=$?" 
tail -15 gpurun_out/bnglue_tests.log
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_trainer_gpu.py tests/test_concat_gpu.py -x -q > gpurun_out/bnglue_decoder.log 2>&1; echo "decoder rc=$?"
tail -15 gpurun_out/bnglue_decoder.log
timeout 900 python tools/bench_decoder.py --config 5 4 3 > gpurun_out/bnglue_bench.jsonl 2> gpurun_out/bnglue_bench.err; echo "bench rc=$?"
cat gpurun_out/bnglue_bench.jsonl; tail -5 gpurun_out/bnglue_bench.err
