#!/bin/bash
# GPU session for the fused head kernels: parity tests, then the head sweep with the TMA ring (default) and register-staged loads (tuning 7=1) at r = 8
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -12 gpurun_out/pytest_gpu.log
python tools/sweep_head.py > gpurun_out/sweep_head_split.json 2> gpurun_out/sweep_head_split.err; echo "sweep split exit $?"
python tools/sweep_head.py --tune 7=1 --only-r 8 > gpurun_out/sweep_head_fused_r8.json 2> gpurun_out/sweep_head_fused_r8.err; echo "sweep fused exit $?"
python - <<'PY'
import json
a=json.load(open('gpurun_out/sweep_head_split.json'))['points']
b={ (p['dtype'],p['enc'],p['kernel']):p for p in json.load(open('gpurun_out/sweep_head_fused_r8.json'))['points']}
for p in a:
    q=b.get((p['dtype'],p['enc'],p['kernel']))
    print("%-5s %-12s %-18s %7.2f us %.3f  %-44s %s" % (p['dtype'],p['enc'],p['kernel'],p['us'],p['frac'],p['variant'], ("| reg-staged %7.2f us %.3f" % (q['us'],q['frac'])) if q else ""))
PY
