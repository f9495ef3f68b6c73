#!/bin/bash
# trainer / optimizer tests and the training configs on one GPU
mkdir -p gpurun_out
python -m pytest tests/test_trainer_gpu.py tests/test_optim_gpu.py tests/test_upsample_gpu.py -q -m gpu --timeout 900 > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_train.log | cut -c1-300
python tools/bench_decoder.py --config 5 4 3 --steps 5 --warmup 3 > gpurun_out/dec_n1.jsonl 2> gpurun_out/dec_n1.err; echo "decoder exit $?"; cat gpurun_out/dec_n1.jsonl | cut -c1-1200; tail -5 gpurun_out/dec_n1.err | cut -c1-400
