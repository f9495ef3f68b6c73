python -m pytest tests/test_head_gpu.py tests/test_guard_bands_gpu.py -m gpu -q -x --timeout 900 2>&1 | tail -2
python tools/sweep_head.py --only-r 8 > gpurun_out/sweep_head_r8.json 2>/dev/null
python - <<'PY'
import json
for p in json.load(open('gpurun_out/sweep_head_r8.json'))['points']:
    print("%-5s %-12s %-18s %7.2f us %.3f  %s" % (p['dtype'],p['enc'],p['kernel'],p['us'],p['frac'],p['variant']))
PY
R=8 bash tools/gpu_ncu_head.sh
