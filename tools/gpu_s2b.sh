#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_daspp_train_gpu.py tests/test_decoder_gpu.py tests/test_trainer_gpu.py -x -q > gpurun_out/s2b_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/s2b_tests.log
timeout 900 python tools/bench_decoder.py --config 5 4 3 > gpurun_out/s2b_bench.jsonl 2> gpurun_out/s2b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/s2b_bench.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config'], d['value'], d['ms_per_step'])
PY
tail -3 gpurun_out/s2b_bench.err
