#!/bin/bash
# GPU session for the decoder-tail kernels: parity tests, smoke, micro-benchmark (f32 + bf16)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
python tools/bench_tail.py > gpurun_out/tail_f32.json 2> gpurun_out/tail_f32.err; echo "tail f32 exit $?"; cat gpurun_out/tail_f32.json; tail -5 gpurun_out/tail_f32.err
python tools/bench_tail.py --dtype bf16 > gpurun_out/tail_bf16.json 2> gpurun_out/tail_bf16.err; echo "tail bf16 exit $?"; cat gpurun_out/tail_bf16.json
python tools/bench_tail.py --batch 32 --height 352 --width 1216 --skip-cpu > gpurun_out/tail_f32_kitti.json 2>/dev/null; cat gpurun_out/tail_f32_kitti.json
