#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_depthconv_gpu.py tests/test_tail_gpu.py tests/test_guard_bands_gpu.py tests/test_decoder_gpu.py tests/test_host_io_gpu.py tests/test_iconv_gpu.py tests/test_lpg_gpu.py -q -m gpu --timeout 900 > gpurun_out/pytest_tail.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_tail.log | cut -c1-300
python tools/bench_tail.py --skip-cpu --skip-literal > gpurun_out/tail_f32.json 2> gpurun_out/tail_f32.err; echo "tail exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/tail_f32.json"))
for k, v in d.items():
    if isinstance(v, dict) and "us" in v:
        print(k, v["us"], v["frac_of_peak"], v.get("kernel"))
PY
