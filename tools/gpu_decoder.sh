#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_decoder_gpu.py -q --timeout 900 > gpurun_out/pytest_decoder.log 2>&1; echo "pytest decoder exit $?"; tail -30 gpurun_out/pytest_decoder.log
timeout 900 python tools/bench_decoder.py --steps 5 --warmup 2 > gpurun_out/decoder_fused_n1.jsonl 2> gpurun_out/decoder_fused_n1.err; echo "decoder fused exit $?"; cat gpurun_out/decoder_fused_n1.jsonl; tail -5 gpurun_out/decoder_fused_n1.err
timeout 900 python tools/bench_decoder.py --steps 3 --warmup 2 --lpg literal > gpurun_out/decoder_literal_n1.jsonl 2> gpurun_out/decoder_literal_n1.err; echo "decoder literal exit $?"; cat gpurun_out/decoder_literal_n1.jsonl; tail -5 gpurun_out/decoder_literal_n1.err
