#!/bin/bash
# quick session after a change to the training path: depthconv / decoder / guard-band tests, decoder benches (configs 3-5)
mkdir -p gpurun_out
python -m pytest tests/test_depthconv_gpu.py tests/test_decoder_gpu.py tests/test_guard_bands_gpu.py -k "depthconv or decoder or depth_conv" -m gpu -q --timeout 600 > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_train.log
timeout 800 python tools/bench_decoder.py --steps 5 --warmup 3 > gpurun_out/decoder_fused_n1.jsonl 2> gpurun_out/decoder_fused_n1.err; echo "decoder exit $?"; cut -c1-330 gpurun_out/decoder_fused_n1.jsonl
