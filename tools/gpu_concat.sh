#!/bin/bash
# GPU session: all GPU tests (incl. concat + decoder), tail/concat micro-benchmark, decoder bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -15 gpurun_out/pytest_gpu.log
python tools/bench_tail.py --skip-cpu > gpurun_out/tail_f32.json 2> gpurun_out/tail_f32.err; echo "tail f32 exit $?"; cat gpurun_out/tail_f32.json; tail -3 gpurun_out/tail_f32.err
python tools/bench_tail.py --skip-cpu --dtype bf16 > gpurun_out/tail_bf16.json 2> gpurun_out/tail_bf16.err; echo "tail bf16 exit $?"; cat gpurun_out/tail_bf16.json
timeout 900 python tools/bench_decoder.py --steps 5 --warmup 2 > gpurun_out/decoder_fused_n1.jsonl 2> gpurun_out/decoder_fused_n1.err; echo "decoder exit $?"; cut -c1-330 gpurun_out/decoder_fused_n1.jsonl; tail -3 gpurun_out/decoder_fused_n1.err
