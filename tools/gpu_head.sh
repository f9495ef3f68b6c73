#!/bin/bash
# fused head kernels: parity tests, then the timing sweep (all scales, both encoders, both dtypes)
mkdir -p gpurun_out
python -m pytest tests/test_head_gpu.py tests/test_guard_bands_gpu.py tests/test_decoder_gpu.py -q -m gpu --timeout 900 > gpurun_out/pytest_head.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_head.log | cut -c1-300
python tools/sweep_head.py > gpurun_out/sweep_head_r02.json 2> gpurun_out/sweep_head.err; echo "sweep exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/sweep_head_r02.json"))
for p in d["points"]:
    print(p["dtype"], p["enc"], p["kernel"], p["variant"], p["us"], p["frac"])
PY
