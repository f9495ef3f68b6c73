#!/bin/bash
# GPU session for the fused head kernels: parity tests (heads, guard bands, decoder), then the head sweep
mkdir -p gpurun_out
python -m pytest tests/test_head_gpu.py tests/test_guard_bands_gpu.py tests/test_decoder_gpu.py -m gpu -q -x --timeout 900 > gpurun_out/pytest_head.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -6 gpurun_out/pytest_head.log
python tools/sweep_head.py > gpurun_out/sweep_head.json 2> gpurun_out/sweep_head.err; echo "sweep exit $?"
python - <<'PY'
import json
for p in json.load(open('gpurun_out/sweep_head.json'))['points']:
    print("%-5s %-12s %-18s %7.2f us %.3f  %s" % (p['dtype'],p['enc'],p['kernel'],p['us'],p['frac'],p['variant']))
PY
