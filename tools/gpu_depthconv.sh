#!/bin/bash
# GPU session for the fused last-convolution forward: parity tests, micro-benchmark (+ register-cap variant), one ncu capture, decoder config 3
mkdir -p gpurun_out
python -m pytest tests/test_depthconv_gpu.py tests/test_decoder_gpu.py "tests/test_guard_bands_gpu.py" -k "depthconv or decoder" -m gpu -q --timeout 600 > gpurun_out/pytest_dcf.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_dcf.log
python tools/bench_tail.py --only-depthconv > gpurun_out/dcf_f32.json 2> gpurun_out/dcf_f32.err; rc=$?; echo "dcf f32 exit $rc"; cat gpurun_out/dcf_f32.json; tail -3 gpurun_out/dcf_f32.err
python tools/bench_tail.py --only-depthconv --dtype bf16 > gpurun_out/dcf_bf16.json 2> gpurun_out/dcf_bf16.err; echo "dcf bf16 exit $?"; cat gpurun_out/dcf_bf16.json
for v in minb3; do
  if [ -f bts-fully-tf_b200/lib/libbtslpg_$v.so ]; then
    BTSLPG_LIB=$PWD/bts-fully-tf_b200/lib/libbtslpg_$v.so python tools/bench_tail.py --only-depthconv --skip-literal > gpurun_out/dcf_f32_$v.json 2> gpurun_out/dcf_f32_$v.err; echo "variant $v exit $?"; cat gpurun_out/dcf_f32_$v.json
  fi
done
if [ $rc -eq 0 ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:depthconv_fwd_mma -c 1 -f -o gpurun_out/dcf_f32 python tools/bench_tail.py --only-depthconv --no-graph --skip-literal --steps 8 > gpurun_out/ncu_dcf.log 2>&1; echo "ncu exit $?"
fi
timeout 600 python tools/bench_decoder.py --config 3 --steps 5 --warmup 3 > gpurun_out/decoder_cfg3.jsonl 2> gpurun_out/decoder_cfg3.err; echo "decoder exit $?"; cat gpurun_out/decoder_cfg3.jsonl | cut -c1-400
